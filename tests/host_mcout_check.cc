// tests/host_mcout_check.cc -- CPU check of the C++ MCout mirror (mcpar_b200/host/mcout.cc):
// same observable behaviour as the reference's MCout (src/mcout.cc:30-145) for add / output /
// collect / maxlike / rewind.  Built and run by tests/test_host_cpp.py.
#include <iostream>
#include <sstream>
#include <string>
#include <string.h>
#include "../mcpar_b200/host/mcout.hh"

#define CHECK(c) do { if (!(c)) { std::cerr << "FAILED: " #c " (line " << __LINE__ << ")\n"; return 1; } } while (0)

int main()
{
  std::ostringstream os;
  MCout out(2, &os, 0);
  CHECK(out.ncol() == 3 && out.size() == 0 && out.nparam() == 2);
  out.newsamps(4);
  CHECK(out.maxsize() == 4 && out.vsize() == 12);
  const Real a[2] = {1.5, -2.25}, b[2] = {0.125, 3.0}, c[2] = {7.0, 8.0};
  out.add(a, -3.5); out.add(b, -0.75);
  CHECK(out.size() == 2 && out.getlval(1) == -0.75 && out.getpset(1)[1] == 3.0);
  out.output();                                    // rows since the last output, "v  v  v  \n"
  CHECK(os.str() == "1.5  -2.25  -3.5  \n0.125  3  -0.75  \n");
  out.output();                                    // nothing new: prints nothing
  CHECK(os.str() == "1.5  -2.25  -3.5  \n0.125  3  -0.75  \n");
  out.add(c, -0.75);                               // tie with row 1: the first maximum is kept (strict >)
  size_t n = 0;
  Real *buf = out.collect(&n);
  CHECK(n == 3 && buf && buf[0] == 7.0 && buf[2] == -0.75);
  delete[] buf;
  buf = out.collect(&n);
  CHECK(n == 0 && buf == 0);
  Real lmax;
  const std::vector<Real> &pm = out.maxlike(&lmax);
  CHECK(lmax == -0.75 && pm[0] == 0.125 && pm[1] == 3.0);
  out.rewind();
  os.str("");
  out.output();
  CHECK(os.str() == "1.5  -2.25  -3.5  \n0.125  3  -0.75  \n7  8  -0.75  \n");
  const Real rows[6] = {1, 2, 9.0, 3, 4, -1.0};
  out.newsamps(2); out.addrows(rows, 2);
  CHECK(out.size() == 5 && out.maxlike(&lmax)[0] == 1.0 && lmax == 9.0);
  // default ostream precision: 6 significant digits, as the reference prints
  std::ostringstream os2; MCout o2(1, &os2, 0); o2.newsamps(1);
  const Real p[1] = {3.14159265358979}; o2.add(p, -1234567.891); o2.output();
  CHECK(os2.str() == "3.14159  -1.23457e+06  \n");
  // binary format: header once, then raw rows in the same order, across several output() calls
  std::ostringstream ob; MCout o3(2, &ob, 0); o3.set_format(MCout::BINARY); o3.newsamps(3);
  o3.add(a, -3.5); o3.output(); o3.add(b, -0.75); o3.add(c, 2.0); o3.output(); o3.output();
  const std::string bin = ob.str();
  CHECK(bin.size() == 16 + 9 * sizeof(Real) && bin.substr(0, 8) == "MCOUTB01");
  int hdr[2]; memcpy(hdr, bin.data() + 8, 8);
  CHECK(hdr[0] == 3 && hdr[1] == (int)sizeof(Real));
  Real back[9]; memcpy(back, bin.data() + 16, sizeof back);
  CHECK(back[0] == 1.5 && back[1] == -2.25 && back[2] == -3.5 && back[3] == 0.125 && back[8] == 2.0);
  std::cout << "ok\n";
  return 0;
}
