"""The closed form of MeshBVH's in-place two-pointer partition (MeshBVH.cs:511-530) that csrc/bvh_device.cuh evaluates in
parallel, against the loop itself: every flag pattern up to 12 items, random patterns up to 3000."""
import itertools
import random


def sequential(flags):
    """The reference's loop on (item, is_right) pairs; returns the permuted item list and mid."""
    arr = list(range(len(flags)))
    i0, i1 = 0, len(arr) - 1
    while i0 <= i1:
        if not flags[arr[i0]]:
            i0 += 1
        else:
            arr[i0], arr[i1] = arr[i1], arr[i0]
            i1 -= 1
    return arr, i0


def closed_form(flags):
    """bvh_device.cuh, db_build_kernel: destination of every position from prefix counts only."""
    n = len(flags)
    rpos = [i for i in range(n) if flags[i]]
    lpos = [i for i in range(n) if not flags[i]]        # front order; l_k = lpos[nL - 1 - k]
    nR, nL = len(rpos), len(lpos)
    if nL == 0 or nL == n:
        return None, nL
    K = sum(1 for k in range(min(nR, nL)) if rpos[k] < lpos[nL - 1 - k])
    lKm1 = n if K == 0 else lpos[nL - K]
    X = rpos[K] if (K < nR and rpos[K] < lKm1) else lKm1 - 1
    out = [None] * n
    kR = 0
    for i in range(n):
        if i <= X:
            dest = ((n if kR == 0 else lpos[nL - kR]) - 1) if flags[i] else i
        else:
            dest = i - 1 if flags[i] else rpos[nL - 1 - (i - kR)]
        assert out[dest] is None, (flags, i, dest)
        out[dest] = i
        kR += 1 if flags[i] else 0
    return out, nL


def check(flags):
    want, mid = sequential(flags)
    got, nL = closed_form(flags)
    assert mid == nL == flags.count(False)
    if got is not None:
        assert got == want, (flags, got, want)


def test_every_pattern_up_to_twelve_items():
    for n in range(1, 13):
        for flags in itertools.product((False, True), repeat=n):
            check(list(flags))


def test_random_patterns():
    rnd = random.Random(5)
    for _ in range(400):
        n = rnd.randrange(13, 3000)
        p = rnd.random()
        check([rnd.random() < p for _ in range(n)])
    for n in (1000, 1001):  # runs: sorted either way, alternating
        check([i >= n // 3 for i in range(n)])
        check([i < n // 3 for i in range(n)])
        check([bool(i & 1) for i in range(n)])


def test_device_introsort_equals_the_host_introsort(tmp_path):
    """The Array.Sort fallback of the device builder (DbIntroSort, one thread) against the host builder's NetIntroSort --
    itself pinned against a Python transcription of .NET's introsort in tests/test_bvh_builder_literal.py -- on the CPU:
    the same permutation, ties and NaNs included (tests/db_introsort_check.cu, host code compiled by nvcc)."""
    import os
    import shutil
    import subprocess
    import pytest
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "db_introsort_check")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, os.path.join(here, "db_introsort_check.cu")], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "same permutation -> ok" in r.stdout, r.stdout[-1000:]
