"""CPU tests of mcpar_b200/mcout_io.py: the text / binary readers against the C++ MCout writer, and the
port of the reference analysis script's iteration bookkeeping (src/anly/mcpar-analysis.R:80-120)."""
import os
import subprocess
import numpy as np

from mcpar_b200 import mcout_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mcpar_b200", "host")

WRITER = r'''
#include <fstream>
#include <iostream>
#include "mcout.hh"
int main(int argc, char **argv) {
  std::ofstream f(argv[1], std::ios::binary);
  MCout t(2, &std::cout, 0), b(2, &f, 0);
  b.set_format(MCout::BINARY);
  t.newsamps(5); b.newsamps(5);
  for (int i = 0; i < 5; ++i) {
    const Real p[2] = {i + 0.25, -1.0 / (i + 1)};
    t.add(p, -0.5 * i); b.add(p, -0.5 * i);
    if (i == 2) { t.output(); b.output(); }
  }
  t.output(); b.output();
  return 0;
}
'''


def test_readers_round_trip(tmp_path):
    src = tmp_path / "w.cc"; src.write_text(WRITER)
    exe = str(tmp_path / "w")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O1", "-I", HOST, "-o", exe, str(src), os.path.join(HOST, "mcout.cc")])
    out = subprocess.check_output([exe, str(tmp_path / "o.bin")]).decode()
    exp = np.array([[i + 0.25, -1.0 / (i + 1), -0.5 * i] for i in range(5)])
    assert np.allclose(mcout_io.read_text(out.splitlines()), exp, rtol=1e-5)
    assert np.array_equal(mcout_io.read_binary(str(tmp_path / "o.bin")), exp)
    assert np.allclose(mcout_io.read_text(["nsamp = 5"] + out.splitlines() + ["max likelihood value: 0"]), exp, rtol=1e-5)


def test_itercount_matches_the_output_order():
    """Rows are dumped in batches of `outstep` iterations; a batch is nproc rank blocks; a rank block is
    iteration-major, then chain (src/mcpar.cc:110-119, src/mcout.cc:52-94)."""
    niter, nproc, npset = 100, 3, 4
    outstep = mcout_io.outstep_of(niter)
    assert outstep == 10
    order = []                                            # the order MCPar::run feeds MCout
    for b in range(niter // outstep):
        for r in range(nproc):
            for t in range(b * outstep, (b + 1) * outstep):
                order += [t + 1] * npset
    assert np.array_equal(mcout_io.itercount(niter, nproc, npset), order)
    assert np.array_equal(mcout_io.itercount_exact(niter, nproc, npset), order)
    # niter not a multiple of outstep (short last batch), and the small-run cadence of 5
    for niter in (64, 23, 5):
        outstep = mcout_io.outstep_of(niter)
        order = []
        for s0 in range(0, niter, outstep):
            for r in range(nproc):
                for t in range(s0, min(niter, s0 + outstep)):
                    order += [t + 1] * npset
        assert np.array_equal(mcout_io.itercount_exact(niter, nproc, npset), order)
