"""CPU ORACLE (test infrastructure) for the COLUMN-SHARDED form of the iteration.

Each rank owns columns [col0, col0 + n_k) of the slack-form A and the matching slices of every
n-vector; b, y, M and all scalars are replicated.  `comm` supplies the only exchanges the path has:
    allreduce_sum(array)   M = sum_k A_k D_k A_k^T, every A.w product, every dot over n
    allreduce_min(scalar)  the ratio test
It exists to check, on the CPU with gloo (world_size 2), that this decomposition reproduces the
unsharded oracle -- i.e. that the GPU path's collectives are the right ones.  It reuses the scalar
logic of ipm_oracle and follows the same reference lines.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import lapack

from . import ipm_oracle as o


class LocalComm:
    """world_size 1."""

    def allreduce_sum(self, a):
        return a

    def allreduce_min(self, v):
        return v


class TorchComm:
    """torch.distributed (gloo on CPU) as the exchange layer."""

    def __init__(self, dist):
        self.dist = dist

    def allreduce_sum(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64))).copy())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        out = t.numpy()
        return out if np.ndim(a) else float(out[0])

    def allreduce_min(self, v):
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t[0])


def solve_sharded(A_k, b, c_k, n_global, comm, tol=1e-8, ip=True, alpha0=0.99995, max_iter=1000):
    """solve_normal_form (interior_point/mod.rs:199-240) on one column shard.  Returns
    (status, x_k / tau, iterations)."""
    m, n_k = A_k.shape
    x = np.ones(n_k)
    z = np.ones(n_k)
    y = np.zeros(m)
    tau = kappa = 1.0

    def residual_scalars():
        rP = b * tau - comm.allreduce_sum(A_k.dot(x))            # feasible_point.rs:122
        rD = c_k * tau - A_k.T.dot(y) - z                        # :123 (local)
        s = comm.allreduce_sum(np.array([rD.dot(rD), c_k.dot(x), x.dot(z)]))
        return rP, rD, math.sqrt(rP.dot(rP)), math.sqrt(s[0]), s[1], float(b.dot(y)), s[2]

    rP, rD, nrp, nrd, cx, by, xz = residual_scalars()
    ini = (nrp, nrd, abs(kappa + cx - by), (xz + tau * kappa) / (n_global + 1))
    for iteration in range(1, max_iter + 1):
        gamma = 1.0 if ip else 0.0
        eta = 1.0 if ip else 1.0 - gamma
        r_G = cx - by + kappa
        mu = (xz + tau * kappa) / (n_global + 1)
        Dinv = x / z
        M = comm.allreduce_sum(A_k.dot(Dinv[:, None] * A_k.T))   # newton_equations.rs:54-57 + all-reduce
        cfac, info = lapack.dpotrf(M, lower=0)
        if info != 0:
            return "NumericalProblem", x / tau, iteration

        def sym_solve(r1, r2):                                   # newton_equations.rs:214-225
            r = r2 + comm.allreduce_sum(A_k.dot(Dinv * r1))
            v, _ = lapack.dpotrs(cfac, r, lower=0)
            return Dinv * (A_k.T.dot(v) - r1), v

        p, q = sym_solve(c_k, b)
        cp_bq = comm.allreduce_sum(np.array([c_k.dot(p)]))[0], float(b.dot(q))

        def delta(xs, tk, g_hat):
            u, v = sym_solve(rD * eta - xs / x, rP * eta)
            cu = comm.allreduce_sum(np.array([c_k.dot(u)]))[0]
            d_tau = (g_hat + 1.0 / tau * tk - (-cu + float(b.dot(v)))) / (1.0 / tau * kappa + (-cp_bq[0] + cp_bq[1]))
            dx = u + p * d_tau
            dz = (xs - z * dx) / x
            return dx, v + q * d_tau, dz, d_tau, 1.0 / tau * (tk - kappa * d_tau)

        def step(dx, dz, d_tau, d_kappa, a0):
            ax = min([1.0] + list((x[dx < 0] / -dx[dx < 0])))
            az = min([1.0] + list((z[dz < 0] / -dz[dz < 0])))
            ax, az = comm.allreduce_min(ax), comm.allreduce_min(az)
            at = min(1.0, tau / -d_tau) if d_tau < 0 else 1.0
            ak = min(1.0, kappa / -d_kappa) if d_kappa < 0 else 1.0
            return min(1.0, ax, at, az, ak) * a0

        xs = (x * -1.0) * z + gamma * mu
        dx, dy, dz, d_tau, d_kappa = delta(xs, gamma * mu - tau * kappa, r_G * eta)
        alpha = step(dx, dz, d_tau, d_kappa, 1.0)
        gamma = o.update_gamma(ip, alpha)
        eta = 1.0 if ip else 1.0 - gamma
        if ip:
            xs = (x * -1.0) * z - (dx * dz) * alpha * alpha + (1.0 - alpha) * gamma * mu
            tk = (1.0 - alpha) * gamma * mu - tau * kappa - alpha * alpha * d_tau * d_kappa
        else:
            xs = (x * -1.0) * z + gamma * mu - dx * dz
            tk = gamma * mu - tau * kappa - d_tau * d_kappa
        dx, dy, dz, d_tau, d_kappa = delta(xs, tk, r_G * eta)
        alpha = 1.0 if ip else step(dx, dz, d_tau, d_kappa, alpha0)
        x, y, z = x + dx * alpha, y + dy * alpha, z + dz * alpha
        tau, kappa = tau + d_tau * alpha, kappa + d_kappa * alpha
        if ip:
            x, z, tau, kappa = np.maximum(x, 1.0), np.maximum(z, 1.0), max(tau, 1.0), max(kappa, 1.0)
        ip = False
        rP, rD, nrp, nrd, cx, by, xz = residual_scalars()
        ind = o.Indicators(rho_p=nrp / max(ini[0], 1.0), rho_d=nrd / max(ini[1], 1.0),
                           rho_A=abs(cx - by) / (tau + abs(by)), rho_g=abs(kappa + cx - by) / max(ini[2], 1.0),
                           rho_mu=((xz + tau * kappa) / (n_global + 1)) / ini[3], obj=cx / tau, bty=by)
        st = ind.status(tau, kappa, tol)
        if st != "Unfinished":
            return st, x / tau, iteration
    return "IterationLimitExceeded", x / tau, max_iter
