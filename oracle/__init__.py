"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU checkers for the B200 MH engine:
  oracle.ref : the reference's own unmodified sources built against shim MPI/MKL
               headers (oracle/_ref/libmcpar_ref{64,32}.so, see oracle/Makefile)
  oracle.mh  : the plain-C restatement of the reference algorithm (mh_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (mcpar_b200, libmcgpu.so) never does.
"""
