// mcutil.cc -- mcutil::qriguess on the GPU (replaces the MKL Sobol stream of src/mcutil.cc).
#include "mcutil.hh"
#include "../../include/mcgpu.h"
#include <stdio.h>
#include <stdlib.h>

void mcutil::qriguess(int rank, int npset, int nparam, const Real plo[], const Real phi[], Real *restrict pout)
{
  const int rc = mcgpu_qriguess(device, rank, npset, nparam, plo, phi, pout);
  if (rc != MCGPU_OK) { fprintf(stderr, "mcutil::qriguess failed (%d): %s\n", rc, mcgpu_last_error(0)); abort(); }
}
