"""The header-only C++ adapters (gtsam_points_b200/cpp) compile against the C ABI; on a GPU box the compiled driver is run
against oracle results in the reference optimizer's call order."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_adapters.cpp")
LIBDIR = os.path.join(ROOT, "gtsam_points_b200", "lib")


def build_driver(tmp_path):
    exe = str(tmp_path / "test_adapters")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "gtsam_points_b200", "cpp", "include"), SRC, "-o", exe, "-L", LIBDIR, "-lb2points", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.check_call(cmd)
    return exe


def test_adapters_compile_against_c_abi(tmp_path):
    from gtsam_points_b200 import build

    build.build()
    assert os.path.exists(build_driver(tmp_path))


@pytest.mark.gpu
def test_adapters_match_oracle_in_reference_call_order(tmp_path):
    import oracle_lib as orc
    from gtsam_points_b200 import synthetic as syn

    exe = build_driver(tmp_path)
    tp, tc = syn.make_cloud(30000, stream=1, scale=0.25)
    sp, sc = syn.make_cloud(12000, stream=2, scale=0.25)
    rng = np.random.default_rng(2)
    Tt = syn.random_pose(rng, 0.4, 3.0)
    Ts = Tt @ syn.random_pose(rng, 0.01, 0.1)
    Ts2 = Tt @ syn.random_pose(rng, 0.01, 0.1)

    def pad4(p):
        return np.concatenate([p, np.ones((len(p), 1))], 1)

    def pad44(c):  # the reference's column-major Matrix4d with zero row/col 3 (symmetric => layout-agnostic)
        m = np.zeros((len(c), 4, 4))
        m[:, :3, :3] = c
        return m

    otgt, osrc = orc.Cloud(tp, tc), orc.Cloud(sp, sc)
    vm = orc.VoxelMap(0.5)
    vm.insert(otgt)
    fv = orc.Factor(vm, osrc, num_threads=4)
    fg = orc.Factor(otgt, osrc, tree=orc.KdTree(otgt, 4), num_threads=4)
    d, d2 = orc.calc_delta(Tt, Ts), orc.calc_delta(Tt, Ts2)
    ev, eg = fv.linearize_raw(d), fg.linearize_raw(d)
    errs = np.array([fv.error(d2), fg.error(d2)])
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        for a in (pad4(tp), pad44(tc), pad4(sp), pad44(sc), Tt, Ts, Ts2, ev, eg, errs, np.array([0.5])):
            a = np.ascontiguousarray(a, dtype=np.float64).ravel()
            f.write(np.uint64(a.size).tobytes())
            f.write(a.tobytes())
    r = subprocess.run([exe, str(case)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK")
