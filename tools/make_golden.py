#!/usr/bin/env python
"""Generate tests/golden/* from the REAL reference (oracle/_ref, built from /root/reference).

Run in the build container only (needs /root/reference to have been compiled by
oracle/Makefile).  One subprocess per density because the reference's init() owns a
function-static RNG.  Outputs, per density rho in {0.5, 1.0} (L = 50, LOOP = 100):

  tests/golden/density<rho>.dat   print_results() text after 100 x force_pair -- asserted
                                  byte-identical to /root/reference/ref_data/density<rho>.dat
  tests/golden/ref_<rho>.npz      pn, npairs, sha256 of q / number_of_partners / half list,
                                  first particle, a 4096-atom sample of q and of p after
                                  100 x force_sorted (full float64), and the 10 printed rows.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
NAMES = {0.5: "density0.5.dat", 1.0: "density1.dat"}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def one(density: float) -> None:
    from oracle.ljoracle import Ref, print_results_lines
    ref = Ref(density, 50.0)
    pn = ref.pn
    q = ref.q_xyz()
    nop, ptr, lst = ref.makepair()
    # (1) the published golden: 100 x force_pair, print_results format
    ref.zero_p()
    ref.force("pair", 100)
    text = "\n".join(print_results_lines(ref.p_xyz())) + "\n"
    with open(os.path.join(GOLD, NAMES[density]), "w") as f:
        f.write(text)
    pub = "/root/reference/ref_data/" + NAMES[density]
    if os.path.exists(pub):
        assert open(pub).read() == text, "mismatch with published " + pub
        print("rho=%.1f: print_results identical to %s" % (density, pub))
    p_pair = ref.p_xyz()
    # (2) full-precision sample after 100 x force_sorted
    ref.zero_p()
    ref.force("sorted", 100)
    p_sorted = ref.p_xyz()
    # (3) the real libstdc++ std::shuffle per row, mt19937(10) (cuda/force_cuda.cu:255-263)
    shuf = lst.copy()
    ref.std_shuffle_rows(shuf, nop, ptr, 10)
    rng = np.random.RandomState(12345)
    sample = np.sort(rng.choice(pn, 4096, replace=False)).astype(np.int64)
    sample[:5] = np.arange(5)
    sample[-5:] = np.arange(pn - 5, pn)
    sample = np.unique(sample)
    np.savez_compressed(
        os.path.join(GOLD, "ref_%.1f.npz" % density),
        density=density, L=50.0, pn=pn, npairs_half=len(lst), steps=100,
        q_sha256=sha(q), nop_half_sha256=sha(nop), list_half_sha256=sha(lst), list_half_shuffled_sha256=sha(shuf),
        nop_half_minmax=np.array([nop.min(), nop.max()]),
        q0=q[0], sample_idx=sample, q_sample=q[sample],
        p_sorted_sample=p_sorted[sample], p_pair_sample=p_pair[sample],
        p_sorted_absmax=np.abs(p_sorted).max(), p_sorted_sum=p_sorted.sum(axis=0))
    print("rho=%.1f: pn=%d half pairs=%d max|p|=%.6f" % (density, pn, len(lst), np.abs(p_sorted).max()))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(float(sys.argv[1]))
    else:
        os.makedirs(GOLD, exist_ok=True)
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        for d in (0.5, 1.0):
            subprocess.check_call([sys.executable, __file__, str(d)])
