"""Multi-GPU exchange through the C ABI (dpomp_comm_*, dpomp_pf_partial_allgather, dpomp_pf_resample_migrate,
dpomp_mbp_resample_migrate): NCCL inside libdpomp.  The world-size-1 cases run on one GPU; the two-rank cases need two
GPUs (`gpurun --gpus 2`) and are skipped otherwise.  Bar: N ranks reproduce one rank bit for bit (random streams are keyed
by GLOBAL theta-particle / chain ids, all host decisions are replicated)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, load_case

pytestmark = pytest.mark.gpu


class _World1:
    """world-size-1 dpomp_comm (no NCCL involved): every exchange entry point degenerates to the local operation."""

    def __init__(self, dp):
        self._h = C.c_void_p()
        dp._capi.check(dp._capi.lib().dpomp_comm_create(None, 0, 0, 1, -1, C.byref(self._h)))
        self.dp = dp

    @property
    def handle(self):
        return self._h

    def __del__(self):
        self.dp._capi.lib().dpomp_comm_destroy(self._h)


def test_world_size_one_comm_is_the_local_operation(dp):
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    dm = dp.device_model(hmm)
    comm = _World1(dp)
    r, w = C.c_int32(-1), C.c_int32(-1)
    dp._capi.check(dp._capi.lib().dpomp_comm_info(comm.handle, C.byref(r), C.byref(w)))
    assert (r.value, w.value) == (0, 1)
    dp._capi.check(dp._capi.lib().dpomp_comm_barrier(comm.handle))
    nb = 6
    thetas = theta[:, None] * np.linspace(0.8, 1.2, nb)[None, :]
    a, b = dp.ParticleFilter(dm, 700, nb, seed=3), dp.ParticleFilter(dm, 700, nb, seed=3)
    a.set_stream_key(11); b.set_stream_key(11)
    ga = a.partial(thetas, 1, 3)
    gb = b.partial_allgather(comm, thetas, 1, 3, nb)
    assert np.array_equal(ga, gb)
    nidx = np.array([2, 2, 5, 1, 6, 6])
    a.permute(nidx); b.resample_migrate(comm, nidx, nb)
    for p in range(nb):
        assert np.array_equal(a.get_pop(p + 1), b.get_pop(p + 1))
    loc = np.arange(12.0).reshape(6, 2)
    out = np.zeros((6, 2))
    dp._capi.check(dp._capi.lib().dpomp_comm_allgather_f64(comm.handle, dp._capi.ptr(loc), 6, 2, dp._capi.ptr(out)))
    assert np.array_equal(out, loc)
    lo, hi = C.c_int64(), C.c_int64()
    for n, world in ((10, 4), (3, 8), (8192, 8)):
        got = []
        for rk in range(world):
            dp._capi.check(dp._capi.lib().dpomp_partition_bounds(n, world, rk, C.byref(lo), C.byref(hi)))
            got.append((lo.value, hi.value))
            assert got[-1] == dp.partition_bounds(n, world, rk)
        assert got[0][0] == 0 and got[-1][1] == n


def _worker(rank, world, port, out_path, what):
    sys.path.insert(0, ROOT)
    import dpomp_b200 as dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    comm = None
    torch.cuda.set_device(rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        comm = dp.Comm()
        assert comm.handle is not None  # NCCL inside libdpomp, not torch collectives
    out = {}
    if what == "smc2":
        model = dp.generate_model("SIS", [100, 1])
        model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
        y = dp.get_observations(os.path.join(ROOT, "tests", "golden", "pooley.csv"))
        hmm = dp.get_private_model(model, y)
        th0 = model.prior.rand(301, np.random.default_rng(5))  # odd count: ragged partition
        res = dp.run_pibis(hmm, th0, 0.5, True, 1.002, 300, rng=np.random.default_rng(6), seed=7, comm=comm, verbose=False)
        out = dict(bme=res.bme, mu=res.mu, theta=res.theta, w=res.weight)
    elif what == "mbpi":
        model = dp.generate_model("SEIR", [100, 0, 1, 0])
        model.prior = dp.UniformProduct([0, 0, 0], [0.02, 1.0, 0.5])
        y = dp.get_observations(os.path.join(ROOT, "tests", "golden", "seir_c3.csv"))[:25]
        hmm = dp.get_private_model(model, y)
        th0 = model.prior.rand(1001, np.random.default_rng(8))
        res = dp.run_mbp_ibis(hmm, th0, 0.5, 3, False, 1.002, rng=np.random.default_rng(9), seed=10, comm=comm,
                              outer_rs=dp.rs_stratified, verbose=False)
        out = dict(bme=res.bme, mu=res.mu, theta=res.theta, w=res.weight)
    elif what == "pmcmc":
        model = dp.generate_model("SIS", [100, 1])
        model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
        y = dp.get_observations(os.path.join(ROOT, "tests", "golden", "pooley.csv"))
        hmm = dp.get_private_model(model, y)
        th0 = np.tile(np.array([[0.003], [0.1]]), (1, 5)) * np.random.default_rng(2).uniform(0.8, 1.2, (2, 5))
        res = dp.run_pmcmc(hmm, th0, steps=60, adapt_period=30, p=512, seed=11, comm=comm, verbose=False)
        out = dict(theta=res.samples.theta, acc=res.accepted)
    if rank == 0:
        np.savez(out_path, **out)
    if world > 1:
        torch.distributed.destroy_process_group()


@pytest.mark.parametrize("what", ["smc2", "mbpi", "pmcmc"])
def test_two_ranks_over_nccl_equal_one_rank_bitwise(tmp_path, what):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2): the NCCL exchange inside libdpomp")
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(1, port, one, what), nprocs=1, join=True)
    mp.spawn(_worker, args=(2, port, two, what), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), (what, k)
