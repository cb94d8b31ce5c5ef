"""Oracle-backed stand-in for the CUDA trajectory store (dpomp_b200.MbpParticles), used ONLY to test the host-side MBP-MCMC
driver on CPU and as the checker of the GPU chains.  Same duck-typed interface; trajectories live in numpy."""
import numpy as np

from oracle import oracle as orc


class OracleMbp:
    def __init__(self, desc, obs_times, t0_index, n, cap=4096, seed=1):
        self.desc, self.obs_times, self.t0_index, self.n, self.cap = desc, list(obs_times), int(t0_index), int(n), int(cap)
        self.C = desc.n_compartments
        self.ic = np.array([desc.initial_condition[c] for c in range(self.C)], dtype=np.int64)
        self.offset, self.key = 0, None
        self.fc = np.tile(self.ic, (n, 1))
        self.t = np.zeros((n, cap)); self.y = np.zeros((n, cap), dtype=np.int32)
        self.len = np.zeros(n, dtype=np.int64); self.ll = np.zeros((n, 2))
        self.prop = [None] * n

    def set_stream_key(self, key): self.key = int(key)
    def set_batch_offset(self, off): self.offset = int(off)

    def iterate(self, theta, obs_i, fresh):
        theta = np.asarray(theta, dtype=np.float64)
        out = np.zeros(theta.shape[1])
        for p in range(theta.shape[1]):
            if fresh:
                self.fc[p] = self.ic; self.len[p] = 0; self.ll[p] = 0.0
            t_start = (theta[self.t0_index - 1, p] if self.t0_index > 0 else 0.0) if obs_i == 1 else self.obs_times[obs_i - 2]
            out[p], self.len[p] = orc.mbp_iterate(self.desc, theta[:, p], self.fc[p], self.t[p], self.y[p], self.len[p], self.ll[p],
                                                  t_start, obs_i, self.key, self.offset + p)
        return out

    def propose(self, theta_i, theta_f, valid, ymax):
        theta_i = np.asarray(theta_i, dtype=np.float64); theta_f = np.asarray(theta_f, dtype=np.float64)
        out = np.full((theta_i.shape[1], 2), -np.inf)
        for p in range(theta_i.shape[1]):
            self.prop[p] = None
            if not valid[p]:
                continue
            times, types, fc, ll, rc = orc.mbp_propose(self.desc, theta_i[:, p], theta_f[:, p], self.t[p], self.y[p], self.len[p],
                                                       self.cap, ymax, self.key, self.offset + p)
            if rc != 0:
                ll = np.array([-np.inf, ll[1]])
            self.prop[p] = (times, types, fc, ll)
            out[p] = ll
        return out

    def accept(self, slots):
        for s in np.asarray(slots, dtype=np.int64) - 1:
            times, types, fc, ll = self.prop[s]
            self.len[s] = len(times); self.t[s, : len(times)] = times; self.y[s, : len(times)] = types
            self.fc[s] = fc; self.ll[s] = ll
