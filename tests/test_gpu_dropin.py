"""THE drop-in test. The reference's own sources are compiled twice from /root/reference/src (oracle/Makefile):
plain (CPU loops) and in the accelerated configuration ACCELERATE_ENABLED / ARCH_NAME=cuda, where
atom::latRho / latDf / latForce (reference src/atom.cpp:159-161,293-295,320-322) call the eight cuda_* hooks of
arch_cuda/cuda_hooks.cpp -> the C ABI -> the sm_100a kernels. The reference's UNMODIFIED driver sequence
(simulation.cpp:137-145,164-194: Verlet, decide, packers and halo exchanges, inter-atom passes -- all host code of
the reference) then runs on both, and the two must agree: occupancy / ids exactly, rho, df, force to 1e-10."""
import numpy as np
import pytest

from misa_md_b200 import synth
from oracle import ref_py as R
from tests import common as cm

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (R.available() and R.available(hooks=True)),
                                                  reason="oracle/_ref libraries not prebuilt")]


def pair(phase, ratio, sigma, dt=0.001, vacancies=0):
    st = cm.make_state(phase, ratio=ratio, sigma=sigma, vacancies=vacancies)
    cpu = R.World(phase, dt=dt)
    gpu = R.World(phase, dt=dt, hooks=True)
    assert gpu.L.ref_accelerated() == 1 and cpu.L.ref_accelerated() == 0
    arr, _ = synth.scatter_to_sub_box(st, (1, 1, 1), (0, 0, 0))
    cpu.atoms(0)[:] = arr
    gpu.atoms(0)[:] = arr
    return cpu, gpu


def check(cpu, gpu, tol_f=1e-10, tol_x=1e-13):
    own = cpu.owned_slices(0)
    a = cpu.atoms(0).reshape(cpu.shape(0))
    b = gpu.atoms(0).reshape(gpu.shape(0))
    assert np.array_equal(a["type"], b["type"]) and np.array_equal(a["id"], b["id"])
    valid = a[own]["type"] >= 0
    for fld, tol in (("rho", 1e-10), ("df", 1e-10), ("f", tol_f), ("x", tol_x)):
        assert cm.rel_err(b[own][fld][valid], a[own][fld][valid]) < tol, fld
    # ghost positions are the host packers' work on both sides: identical up to the owners' positions
    assert cm.rel_err(b["x"], a["x"]) < tol_x


def test_reference_driver_with_cuda_hooks_matches_cpu_reference():
    cpu, gpu = pair((9, 8, 10), (90, 6, 4), 0.05, vacancies=7)
    cpu.prepare(); gpu.prepare()
    check(cpu, gpu)
    for _ in range(5):
        cpu.step(); gpu.step()
    check(cpu, gpu, tol_f=1e-9, tol_x=1e-12)
    cpu.close(); gpu.close()


def test_reference_driver_with_cuda_hooks_through_a_cascade():
    """Host-side inter-atom code of the reference (interRho / interForce / decide) interleaved with the hooks."""
    cpu, gpu = pair((10, 10, 10), (1, 0, 0), 0.0, dt=2e-4)
    cpu.prepare(); gpu.prepare()
    lat, direction = (5, 5, 5, 0), (1.0, 3.0, 5.0)
    cpu.collision_step(lat, direction, 300.0); gpu.collision_step(lat, direction, 300.0)
    for _ in range(100):
        cpu.step(); gpu.step()
    assert cpu.total_inter() > 0 and cpu.total_inter() == gpu.total_inter()
    assert np.array_equal(cpu.inter(0)["id"], gpu.inter(0)["id"])
    check(cpu, gpu, tol_f=1e-7, tol_x=1e-9)
    cpu.close(); gpu.close()
