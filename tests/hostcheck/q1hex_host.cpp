// Test-only host instantiation of csrc/q1hex_math.cuh (checks the sum-factorised algebra on the CPU).
#include "../../galerkintoolkit.jl_b200/csrc/q1hex_math.cuh"
extern "C" void q1hex_host(const double* X24, double alpha, double fscale, double* Ke36, double* be8) {
  double X[8][3];
  for (int v = 0; v < 8; ++v) for (int k = 0; k < 3; ++k) X[v][k] = X24[v * 3 + k];
  q1hex::Cell<double> g;
  q1hex::geometry<double>(X, g);
  double Ke[36], be[8];
  q1hex::laplace_ke<double>(g, alpha, Ke);
  q1hex::source_be<double>(g, fscale, be);
  for (int i = 0; i < 36; ++i) Ke36[i] = Ke[i];
  for (int i = 0; i < 8; ++i) be8[i] = be[i];
}
