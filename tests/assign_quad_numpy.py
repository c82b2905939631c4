"""A second, independent restatement of estimator.f:284-512 (assign_quad), written straight from the Fortran text as a sequential
Python loop over float32 numpy scalars -- not derived from oracle/estimator_oracle.c.  REAL arithmetic throughout, the cell indices of
f:308-315 / 342-351 (grid A: rows 2*cell-1, the base cell is NOT wrapped; grid B: rows 2*cell, the base cell IS wrapped), the weights of
f:316-320, the products (hx*hy)*hz*we.  Small catalogues only (pure Python)."""
import numpy as np

f32 = np.float32


def _weights(h):
    h2 = f32(h * h)
    omh = f32(f32(1.) - h)
    m2 = f32(f32(omh * omh) * omh)                                 # (1.-h)**3
    m1 = f32(f32(4.) + f32(f32(f32(f32(3.) * h) - f32(6.)) * h2))  # 4.+(3.*h-6.)*h2
    p2 = f32(h2 * h)
    p1 = f32(f32(f32(f32(6.) - m2) - m1) - p2)                     # 6.-m2-m1-p2
    return m2, m1, p1, p2


def assign_quad(r, w, dtl, kf_ks, offset, ia=0, ib=0, ic=0, id=0):
    """r (3,Np) float32, w (Np) float32, dtl (2*Ngrid,Ngrid,Ngrid) float32 indexed [row,iy,iz] (1-based in the Fortran), in place."""
    N = dtl.shape[1]
    kf_ks, offset = f32(kf_ks), f32(offset)
    md = lambda a, b: int(np.fmod(a, b))                           # Fortran MOD: sign of the first argument
    for i in range(r.shape[1]):
        x = [f32(r[0, i]), f32(r[1, i]), f32(r[2, i])]
        if ia == 0 and ib == 0 and ic == 0 and id == 0:
            we = f32(w[i])
        else:
            rn = f32(f32(f32(x[0] * x[0]) + f32(x[1] * x[1])) + f32(x[2] * x[2]))
            if ic == 0 and id == 0:
                we = f32(f32(f32(f32(w[i]) * x[ia - 1]) * x[ib - 1]) / rn)
            else:
                we = f32(f32(f32(f32(f32(f32(w[i]) * x[ia - 1]) * x[ib - 1]) * x[ic - 1]) * x[id - 1]) / f32(rn * rn))
        rows_a, rows_b, wa, wb = [], [], [], []
        for a in range(3):
            rx = f32(f32(f32(kf_ks * x[a]) + f32(1.)) + offset)
            tx = f32(rx + f32(0.5))
            im1 = int(rx)                                          # int() truncates
            cells = [md(im1 - 2 + N, N) + 1, im1, md(im1, N) + 1, md(im1 + 1, N) + 1]        # m2, m1 (not wrapped), p1, p2
            rows_a.append(cells)
            wa.append(_weights(f32(rx - f32(im1))))
            nm1 = int(tx)
            g = f32(tx - f32(nm1))
            nm1 = md(nm1 - 1, N) + 1
            rows_b.append([md(nm1 - 2 + N, N) + 1, nm1, md(nm1, N) + 1, md(nm1 + 1, N) + 1])
            wb.append(_weights(g))
        for kz in range(4):
            for ky in range(4):
                for kx in range(4):
                    v = f32(f32(f32(wa[0][kx] * wa[1][ky]) * wa[2][kz]) * we)
                    dtl[2 * rows_a[0][kx] - 1 - 1, rows_a[1][ky] - 1, rows_a[2][kz] - 1] += v
        for kz in range(4):
            for ky in range(4):
                for kx in range(4):
                    v = f32(f32(f32(wb[0][kx] * wb[1][ky]) * wb[2][kz]) * we)
                    dtl[2 * rows_b[0][kx] - 1, rows_b[1][ky] - 1, rows_b[2][kz] - 1] += v
    return dtl
