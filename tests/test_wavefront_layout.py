"""The systolic wavefront form of the in-place a-trous pass (csrc/wavefront_layout.h, wavefront.cuh) rests on a schedule
argument: with chains in lock step at i = t - 3 r - cx, every filtered value a pixel needs is already in its band's
32-entry history rings, whether the halo warp brings the rows above just in time or 16 steps early.  tests/wf_sim.cpp replays that schedule on the CPU with tagged history entries -- built from the
product's own layout header -- against the literal row-major in-place loop: sizes with odd / even widths and heights,
one-pixel images, row tiles that start and end anywhere."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def wf_sim(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("wf") / "wf_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "wf_sim.cpp")], check=True)
    return exe


CASES = [(80, 48, 0, 48), (33, 16, 0, 16), (1, 1, 0, 1), (2, 9, 0, 9), (5, 7, 0, 7), (3, 3, 0, 3), (8, 8, 0, 8), (7, 33, 0, 33),
         (64, 64, 16, 40), (37, 41, 8, 41), (16, 64, 12, 52), (200, 31, 4, 27), (96, 54, 0, 54), (61, 40, 7, 38), (480, 270, 0, 270),
         (4, 2, 0, 2), (9, 5, 1, 4), (130, 67, 0, 67)]


@pytest.mark.parametrize("w,h,y0,y1", CASES)
def test_schedule_reproduces_the_literal_in_place_pass(wf_sim, w, h, y0, y1):
    for seed, early in ((1, 0), (7, 1), (3, 1)):
        r = subprocess.run([wf_sim, str(w), str(h), str(y0), str(y1), str(seed), str(early)], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        assert "violations 0, missing 0, differing 0 -> ok" in r.stdout


def test_simulator_notices_a_broken_schedule(tmp_path):
    """Negative control: with a lag of 2 instead of 3 the value at (x + 4, y - 2) is not there yet, with 8-entry rings values are
    overwritten while still needed, with 16-entry rings a halo warp that runs 16 steps ahead overwrites them; the simulator
    must report all three."""
    src = open(os.path.join(ROOT, "yetanotherconsolegameengine_b200", "csrc", "wavefront_layout.h")).read()
    for a, b, early in (("#define YCGE_WF_L 3 ", "#define YCGE_WF_L 2 ", 0), ("#define YCGE_WF_RING 32 ", "#define YCGE_WF_RING 8 ", 0),
                        ("#define YCGE_WF_RING 32 ", "#define YCGE_WF_RING 16 ", 1)):
        assert a in src
        d = tmp_path / (b.split()[1] + b.split()[2] + str(early)) / "x"
        (d / "yetanotherconsolegameengine_b200" / "csrc").mkdir(parents=True)
        (d / "tests").mkdir()
        (d / "yetanotherconsolegameengine_b200" / "csrc" / "wavefront_layout.h").write_text(src.replace(a, b))
        (d / "tests" / "wf_sim.cpp").write_text(open(os.path.join(ROOT, "tests", "wf_sim.cpp")).read())
        exe = str(d / "sim")
        subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, str(d / "tests" / "wf_sim.cpp")], check=True)
        r = subprocess.run([exe, "80", "48", "0", "48", "1", str(early)], stdout=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "FAIL" in r.stdout
