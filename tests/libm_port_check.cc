// Test shim: exposes the HOST build of the product's libm port (stage_b200/csrc/rt_libm.cuh) so
// tests/test_libm_port.py can compare it with the C library's sin()/cos() bit for bit.
#include "../stage_b200/csrc/rt_libm.cuh"
extern "C" void port_sincos(unsigned long long n, const double *x, double *s, double *c)
{
  for (unsigned long long i = 0; i < n; i++) {
    s[i] = stg::libm::sin_host_exact(x[i]);
    c[i] = stg::libm::cos_host_exact(x[i]);
  }
}
extern "C" void libm_sincos(unsigned long long n, const double *x, double *s, double *c)
{
  for (unsigned long long i = 0; i < n; i++) {
    s[i] = sin(x[i]);
    c[i] = cos(x[i]);
  }
}
