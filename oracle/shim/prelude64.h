/* oracle/shim/prelude64.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Force-included (g++ -include) in front of the reference's UNMODIFIED sources
 * (/root/reference/src/{mcpar,mcout,rosenbrock,mcutil}.cc) to build them as
 * fp64: every system header the reference pulls in is included first, then
 * `float` is re-spelled `double`.  A bare -Dfloat=double would break libstdc++'s
 * own traits; doing it after the system headers does not.  All float literals in
 * the reference (100.0f, 0.5f, 5.0f, 1.0f ...) are exactly representable, so they
 * promote losslessly; FPEPS=1.0e-14 is a double literal already (mcpar.cc:15).
 */
#ifndef ORACLE_PRELUDE64_H_
#define ORACLE_PRELUDE64_H_
#include <iostream>
#include <fstream>
#include <sstream>
#include <iomanip>
#include <vector>
#include <limits>
#include <exception>
#include <new>
#include <string>
#include <memory>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <chrono>
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <assert.h>
#include <unistd.h>
#define ORACLE_REAL_IS_DOUBLE 1
#define float double
#endif
