"""Oracle-backed stand-in for the CUDA ParticleFilter handle, used ONLY to test the host-side outer layers (sharding,
migration, bookkeeping) on CPU.  Same duck-typed interface as dpomp_b200.ParticleFilter; populations live in numpy."""
import numpy as np
import torch

from oracle import oracle as orc


class OraclePF:
    def __init__(self, desc, n_particles, n_batch, rs_type=1, seed=1):
        self.desc, self.n, self.nb, self.rs_type, self.seed = desc, n_particles, n_batch, rs_type, seed
        self.C = desc.n_compartments
        self.pops = np.zeros((n_batch, n_particles, self.C), dtype=np.int64)
        self.offset, self.ids, self.key, self.calls = 0, None, None, 0
        self.filter_words = self.C * n_particles
        self.n_particles, self.n_batch = n_particles, n_batch

    def set_batch_offset(self, off): self.offset = int(off)
    def set_filter_ids(self, ids): self.ids = None if ids is None else np.asarray(ids, dtype=np.int64)
    def set_stream_key(self, key): self.key = int(key)

    def _key(self):
        self.calls += 1
        k, self.key = self.key, None
        return k if k is not None else (self.seed * 1000003 + self.calls)

    def partial(self, theta, ymin, ymax):
        theta = np.asarray(theta, dtype=np.float64)
        if theta.ndim == 1:
            theta = theta[:, None]
        key = self._key()
        out = np.zeros(theta.shape[1])
        for j in range(theta.shape[1]):
            fid = int(self.ids[j]) if self.ids is not None else self.offset + j
            out[j] = orc.pf_partial(self.desc, theta[:, j], self.n, self.pops[j], ymin, ymax, self.rs_type, key, fid)[0]
        return out

    def loglik(self, theta):
        return self.partial(theta, 1, self.desc.n_obs)

    def permute(self, nidx):
        idx = np.asarray(nidx, dtype=np.int64) - 1
        self.pops[: len(idx)] = self.pops[idx].copy()

    def copy_from(self, src, dst_slots, src_slots):
        d = np.asarray(dst_slots, dtype=np.int64) - 1
        s = np.asarray(src_slots, dtype=np.int64) - 1
        if len(d):
            self.pops[d] = src.pops[s]

    def export_tensor(self, slots):
        s = np.asarray(slots, dtype=np.int64) - 1
        return torch.from_numpy(self.pops[s].astype(np.int32).reshape(-1).copy())

    def import_tensor(self, slots, buf):
        s = np.asarray(slots, dtype=np.int64) - 1
        if len(s):
            self.pops[s] = buf.numpy().reshape(len(s), self.n, self.C).astype(np.int64)
