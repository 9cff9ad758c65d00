"""Pair-list cache files of the reference: the text cache of cuda/force_cuda.cu and
the binary pair.dat of cpu_ref, exchanged with the REAL reference where it is available."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_text_cache_format_roundtrip_and_validation(tmp_path, oracle):
    from lj_gpu_b200 import LJError, loadpair, makepaircache
    q = oracle.init_fcc(0.5, 12.0)
    nop, ptr, lst = oracle.makepair(q, full=True)
    path = str(tmp_path / ".cache_pair_all.dat")
    makepaircache(path, nop, ptr.astype(np.int32), lst)
    # the exact text makepaircache() of the reference writes (cuda/force_cuda.cu:165-176)
    lines = open(path).read().split("\n")
    assert lines[0] == "%d %d" % (len(nop), len(lst))
    assert lines[1] == "%d %d" % (nop[0], ptr[0]) and lines[len(nop)] == "%d %d" % (nop[-1], ptr[-1])
    assert lines[len(nop) + 1] == str(lst[0]) and lines[len(nop) + len(lst)] == str(lst[-1]) and lines[-1] == ""
    a, b, c = loadpair(path, len(nop))
    assert np.array_equal(a, nop) and np.array_equal(b, ptr) and np.array_equal(c, lst)
    with pytest.raises(LJError):                     # header pn must match (force_cuda.cu:210-213)
        loadpair(path, len(nop) + 1)
    bad = lines[:]
    bad[len(nop) + 5] = str(len(nop) + 7)            # sorted_list entry out of range (check_loadedpair)
    open(path, "w").write("\n".join(bad))
    with pytest.raises(LJError):
        loadpair(path)
    with pytest.raises(LJError):
        loadpair(str(tmp_path / "missing.dat"))


_WORKER = r"""
import sys, os, hashlib
import numpy as np
sys.path.insert(0, %(root)r)
os.chdir(%(cwd)r)
from oracle.ljoracle import Ref
from lj_gpu_b200 import loadpair_dat, savepair_dat
ref = Ref(0.5, 16.0)
ref.savepair()                                   # the reference writes ./pair.dat
nop, ip, jp = ref.pair_arrays()
a, b, c = loadpair_dat("pair.dat", ref.pn)       # ... and this library reads it
assert np.array_equal(a, nop) and np.array_equal(b, ip) and np.array_equal(c, jp)
sha_ref = hashlib.sha256(open("pair.dat", "rb").read()).hexdigest()
os.rename("pair.dat", "pair_ref.dat")
savepair_dat("pair.dat", nop, ip, jp)            # this library writes ...
assert os.path.getsize("pair.dat") == 4 + 4 * (400000 + 2 * 12000000)
assert hashlib.sha256(open("pair.dat", "rb").read()).hexdigest() == sha_ref   # byte-identical file
ref.lib.ljref_zero_p()
ref.loadpair()                                   # ... and the reference reads it back
n2, i2, j2 = ref.pair_arrays()
assert np.array_equal(n2, nop) and np.array_equal(i2, ip) and np.array_equal(j2, jp)
print("ok", ref.pn, len(ip))
"""


def test_pair_dat_exchanged_with_the_real_reference(tmp_path):
    from oracle import ljoracle as lo
    if not lo.have_ref(0.5):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = subprocess.run([sys.executable, "-c", _WORKER % dict(root=ROOT, cwd=str(tmp_path))],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.startswith("ok")


def test_pair_dat_roundtrip_without_the_reference(tmp_path, oracle):
    from lj_gpu_b200 import LJError, loadpair_dat, savepair_dat
    q = oracle.init_fcc(1.0, 9.0)
    nop, ptr, lst = oracle.makepair(q, full=False)
    ip = np.repeat(np.arange(len(nop), dtype=np.int32), nop)
    path = str(tmp_path / "pair.dat")
    savepair_dat(path, nop, ip, lst)
    a, b, c = loadpair_dat(path, len(nop))
    assert np.array_equal(a, nop) and np.array_equal(b, ip) and np.array_equal(c, lst)
    with open(path, "r+b") as f:                      # corrupt one j index
        f.seek(4 + 4 * (400000 + 12000000) + 8)
        f.write(np.int32(len(nop) + 1).tobytes())
    with pytest.raises(LJError):
        loadpair_dat(path, len(nop))
