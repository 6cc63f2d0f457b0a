#!/bin/bash
# dev: build a copy of the package with extra nvcc flags -> variants/<name>/ilupp_b200 (only the listed .cu are recompiled)
# usage: mk_variant.sh <name> "<nvcc flags>" file1.cu [file2.cu ...]
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; shift 2
dst=variants/$name/ilupp_b200
mkdir -p variants/$name; ln -sfn ../../include variants/$name/include
rm -rf $dst
cp -rp ilupp_b200 $dst
for f in "$@"; do rm -f $dst/csrc/build/${f%.cu}.o; done
make -C $dst/csrc NVEXTRA="$flags" > variants/$name/build.log 2>&1 || { tail -30 variants/$name/build.log; exit 1; }
grep -A2 "k_trsv_systolicILb0ELb1ELb0" $dst/csrc/build/ilub_systolic.ptxas.log | grep Used || true
