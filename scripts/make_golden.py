"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/_ilupp*.so, compiled from
/root/reference/src/binding.cpp by oracle/Makefile).  Run in the build container only; the fixtures
are committed because /root/reference does not exist on the GPU box.

The cases are the reference's own hot-path test problems (test/tests.py:9-43: laplace(50),
laplace2d(49), random(50, seed 39273)) in CSR and CSC, plus one nonsymmetric tridiagonal matrix
for the apply_trans hole noted in SURVEY.md section 4.
"""
import os
import sys

import numpy as np
import scipy.sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def laplace_matrix(n, format='csr'):  # test/tests.py:9-12
    h = 1.0 / (n + 1)
    d = np.ones(n) / (h ** 2)
    return scipy.sparse.diags((-d[:-1], 2 * d, -d[:-1]), (-1, 0, 1)).asformat(format)


def laplace2d_matrix(n_total, format='csr'):  # test/tests.py:14-17
    n = int(np.sqrt(n_total))
    A, I = laplace_matrix(n), scipy.sparse.eye(n)
    return (scipy.sparse.kron(A, I) + scipy.sparse.kron(I, A)).asformat(format)


def random_matrix(n, format='csr', eye_factor=10.0):  # test/tests.py:33-36
    density = min(1.0, 5 / n)
    return (scipy.sparse.random(n, n, density=density, random_state=39273) +
            eye_factor * scipy.sparse.eye(n)).asformat(format)


def nonsym_tridiag(n, format='csr'):
    rng = np.random.default_rng(7)
    return scipy.sparse.diags((-rng.random(n - 1) - 0.5, 3.0 + rng.random(n), -rng.random(n - 1) - 0.1),
                              (-1, 0, 1)).asformat(format)


PROBLEMS = {"laplace": laplace_matrix, "laplace2d": laplace2d_matrix, "random": random_matrix,
            "nonsym_tridiag": nonsym_tridiag}


def fields(A):
    A = A.copy()
    A.sort_indices()
    return (np.ascontiguousarray(A.data, np.float64), np.ascontiguousarray(A.indices, np.int32),
            np.ascontiguousarray(A.indptr, np.int32))


def main():
    ref = oracle.ref_module()
    if ref is None:
        raise SystemExit("oracle/_ref/_ilupp*.so missing: run `make -C oracle ref` where /root/reference exists")
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, gen in PROBLEMS.items():
        for fmt in ("csr", "csc"):
            A = gen(50, fmt)
            data, idx, ptr = fields(A)
            is_csr = fmt == "csr"
            n = A.shape[0]
            b = np.random.default_rng(11).standard_normal(n)
            rec = {"A_data": data, "A_indices": idx, "A_indptr": ptr, "is_csr": np.array(is_csr), "b": b}
            P = ref.ILU0Preconditioner(data, idx, ptr, is_csr)
            (Ld, Li, Lp, Lc, _, _), (Ud, Ui, Up, Uc, _, _) = P.factors_info()
            y = b.copy(); P.apply(y)
            yt = b.copy(); P.apply_trans(yt)
            rec.update(ilu0_L_data=Ld, ilu0_L_indices=Li, ilu0_L_indptr=Lp, ilu0_L_is_csr=np.array(Lc),
                       ilu0_U_data=Ud, ilu0_U_indices=Ui, ilu0_U_indptr=Up, ilu0_U_is_csr=np.array(Uc),
                       ilu0_apply=y, ilu0_apply_trans=yt, ilu0_total_nnz=np.array(P.total_nnz))
            if name in ("laplace", "laplace2d"):
                As = A
            else:
                As = ((A + A.T) / 2).asformat(fmt)  # symmetrised like tests.py:220-221
            sdata, sidx, sptr = fields(As)
            C = ref.IChol0Preconditioner(sdata, sidx, sptr, is_csr)
            ((Cd, Ci, Cp, Cc, _, _),) = C.factors_info()
            yc = b.copy(); C.apply(yc)
            yct = b.copy(); C.apply_trans(yct)
            rec.update(S_data=sdata, S_indices=sidx, S_indptr=sptr,
                       ichol0_L_data=Cd, ichol0_L_indices=Ci, ichol0_L_indptr=Cp, ichol0_L_is_csr=np.array(Cc),
                       ichol0_apply=yc, ichol0_apply_trans=yct, ichol0_total_nnz=np.array(C.total_nnz))
            np.savez_compressed(os.path.join(out, "%s_%s.npz" % (name, fmt)), **rec)
            print("wrote", name, fmt, "n", n, "nnz", A.nnz)
    # factors + apply of the reference's other triangular preconditioners (their threshold
    # factorisations stay on the CPU; FactorsPreconditioner applies such factors on the GPU)
    for name, gen, sym in (("laplace2d", laplace2d_matrix, True), ("random", random_matrix, False)):
        A = gen(144 if name == "laplace2d" else 120, "csr")
        data, idx, ptr = fields(A)
        n = A.shape[0]
        b = np.random.default_rng(12).standard_normal(n)
        rec = {"A_data": data, "A_indices": idx, "A_indptr": ptr, "is_csr": np.array(True), "b": b}
        kinds = [("ilut", ref.ILUTPreconditioner, (4, 1e-3)), ("iluc", ref.ILUCPreconditioner, (4, 1e-3))]
        if sym:
            kinds.append(("icholt", ref.ICholTPreconditioner, (3, 1e-3)))
        for key, cls, args in kinds:
            Pq = cls(data, idx, ptr, True, *args)
            fi = Pq.factors_info()
            for fname, (Fd, Fi, Fp, Fc, _, _) in zip(("L", "U"), fi):
                rec.update({"%s_%s_data" % (key, fname): Fd, "%s_%s_indices" % (key, fname): Fi,
                            "%s_%s_indptr" % (key, fname): Fp, "%s_%s_is_csr" % (key, fname): np.array(Fc)})
            y = b.copy(); Pq.apply(y)
            yt = b.copy(); Pq.apply_trans(yt)
            rec.update({key + "_apply": y, key + "_apply_trans": yt, key + "_total_nnz": np.array(Pq.total_nnz)})
        np.savez_compressed(os.path.join(out, "factors_%s.npz" % name), **rec)
        print("wrote factors", name, "n", n, [k for k, _, _ in kinds])
    # SURVEY.md 8(f)1, permuted triangular solves (sparse_implementation.h:4166-4258): factors, pivoting permutations and
    # apply / apply_trans of the reference's ILUTP / ILUCP on a matrix with a weakened diagonal (so that pivoting happens)
    for fmt in ("csr", "csc"):
        A = (random_matrix(120, fmt) - 9.5 * scipy.sparse.eye(120)).asformat(fmt)
        data, idx, ptr = fields(A)
        n = A.shape[0]
        b = np.random.default_rng(13).standard_normal(n)
        rec = {"A_data": data, "A_indices": idx, "A_indptr": ptr, "is_csr": np.array(fmt == "csr"), "b": b}
        for key, cls in (("ilutp", ref.ILUTPPreconditioner), ("ilucp", ref.ILUCPPreconditioner)):
            Pq = cls(data, idx, ptr, fmt == "csr", 10, 1e-3, 1.0, -1, 10.0)
            for fname, (Fd, Fi, Fp, Fc, _, _) in zip(("L", "U"), Pq.factors_info()):
                rec.update({"%s_%s_data" % (key, fname): Fd, "%s_%s_indices" % (key, fname): Fi,
                            "%s_%s_indptr" % (key, fname): Fp, "%s_%s_is_csr" % (key, fname): np.array(Fc)})
            pl, pr = Pq.permutations()
            rec[key + "_perm_left"] = np.array([] if pl is None else pl, dtype=np.int32)
            rec[key + "_perm_right"] = np.array([] if pr is None else pr, dtype=np.int32)
            y = b.copy(); Pq.apply(y)
            yt = b.copy(); Pq.apply_trans(yt)
            rec.update({key + "_apply": y, key + "_apply_trans": yt, key + "_total_nnz": np.array(Pq.total_nnz)})
        np.savez_compressed(os.path.join(out, "factors_perm_%s.npz" % fmt), **rec)
        print("wrote permuted factors", fmt)
    # .isp files written by the reference's own matrix_sparse::write_binary (sparse_implementation.h:3839-3860)
    # -- the ILU0 factors of the nonsymmetric tridiagonal problem (CSR and CSC input: ROW- and COLUMN-oriented files)
    g = np.load(os.path.join(out, "nonsym_tridiag_csr.npz"))
    for fmt in ("csr", "csc"):
        A = scipy.sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"])).asformat(fmt)
        R = oracle.ref.ILU0(A)
        R.write_factor_binary(0, os.path.join(out, "isp_ilu0_tridiag_%s-L.isp" % fmt))
        R.write_factor_binary(1, os.path.join(out, "isp_ilu0_tridiag_%s-R.isp" % fmt))
    print("wrote .isp fixtures")
    # BiCGStab with ILU0 through the reference templates (harness)
    from ilupp_b200 import problems
    A = problems.convdiff3d(12)
    b = A @ np.ones(A.shape[0])
    R = oracle.ref.ILU0(A)
    x, it, rel, ab, ok = R.bicgstab(A, b, 1e-8, 1e-8, 2000)
    data, idx, ptr = fields(A)
    np.savez_compressed(os.path.join(out, "bicgstab_convdiff12.npz"), A_data=data, A_indices=idx, A_indptr=ptr, b=b,
                        x=x, iters=np.array(it), rel=np.array(rel), abs=np.array(ab), ok=np.array(ok))
    print("wrote bicgstab_convdiff12: iters", it, "ok", ok)


if __name__ == "__main__":
    main()
