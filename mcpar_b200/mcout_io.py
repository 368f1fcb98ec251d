"""Readers for what the drivers write, and the iteration bookkeeping the reference's analysis script
relies on (host-side helpers; no GPU involved).

  read_text    the reference's text format: one row per line, every column followed by two blanks
               (src/mcout.cc:37-47), optionally behind the `nsamp = N` banner of mcpar-rosen1
  read_binary  the binary format of mcpar_b200/host/mcout.cc (MCout::BINARY): 16-byte header
               "MCOUTB01", int32 ncol, int32 sizeof(Real), then rows of ncol reals, little endian
  itercount    port of mcparam.itercount (src/anly/mcpar-analysis.R:80-120): the iteration number of
               every output row, given that rows are dumped in batches of `outstep` iterations, each
               batch rank-major, each rank block iteration-major then chain (src/mcpar.cc:110-119,
               src/mcout.cc:52-94)
"""
import numpy as np

MAGIC = b"MCOUTB01"


def read_text(path_or_lines):
    lines = open(path_or_lines).read().splitlines() if isinstance(path_or_lines, str) else list(path_or_lines)
    if lines and lines[0].startswith("nsamp = "):
        lines = lines[1:]
    rows = [[float(t) for t in l.split()] for l in lines if l.strip() and not l.startswith("max likelihood")]
    n = len(rows[0]) if rows else 0
    return np.array([r for r in rows if len(r) == n], dtype=np.float64)


def read_binary(path):
    raw = open(path, "rb").read()
    if raw[:8] != MAGIC:
        raise ValueError("not an MCout binary file")
    ncol, size = np.frombuffer(raw[8:16], dtype="<i4")
    dt = {4: "<f4", 8: "<f8"}[int(size)]
    body = np.frombuffer(raw[16:], dtype=dt)
    if body.size % ncol:
        raise ValueError("truncated MCout binary file")
    return body.reshape(-1, int(ncol)).astype(np.float64)


def outstep_of(niter):
    """Output cadence of MCPar::run (src/mcpar.cc:110)."""
    return niter // 10 if niter > 50 else 5


def itercount(niter, nproc, npset):
    """1-based iteration number of every output row (mcpar-analysis.R:80-120).  As in the R code the
    batch index divides by nbatch*nchain, which equals the true batch size outstep*nchain only when
    niter // outstep == outstep (e.g. niter = 100); itercount_exact has the general rule."""
    ntot = niter * nproc * npset
    nchain = nproc * npset
    outstep = outstep_of(niter)
    nbatch = niter // outstep
    slot = np.arange(ntot)
    out_batch = slot // (nbatch * nchain)
    nblock = outstep * npset
    seq_block = 1 + (np.arange(nblock) // npset)
    return out_batch * outstep + np.resize(seq_block, ntot)


def itercount_exact(niter, nproc, npset):
    """The same bookkeeping for any niter: batches of min(outstep, remaining) iterations."""
    outstep = outstep_of(niter)
    out = []
    done = 0
    while done < niter:
        n = min(outstep, niter - done)
        block = done + 1 + (np.arange(n * npset) // npset)
        out.append(np.tile(block, nproc))
        done += n
    return np.concatenate(out)
