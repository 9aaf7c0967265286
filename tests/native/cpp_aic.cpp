// TEST INFRASTRUCTURE: exercises the C++ mirror include/volcanor_b200.hpp like the reference's own unit tests do
// (tests/wing1x3_test.f90:80-98: rotor%calcAIC; then gamVec = AIC_inv*RHS).
//   cpp_aic <wingpanel-records.bin> <nc> <ns>        prints AIC (row-major rows) and the solution for RHS = 1..N
//   cpp_aic --solve-before-calcAIC                   must fail like the reference aborts: message + exit code 3
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/volcanor_b200.hpp"

int main(int argc, char** argv) {
  try {
    vlc::Context ctx(0);
    if (argc == 2 && std::strcmp(argv[1], "--solve-before-calcAIC") == 0) {
      vlc::Rotor rotor(ctx, 0, 1, 1, 2, 2, 0);
      double rhs[2] = {1.0, 1.0}, g[2];
      rotor.solve(rhs, g);
      std::puts("unexpected success");
      return 1;
    }
    if (argc != 4) return 2;
    const int nc = std::atoi(argv[2]), ns = std::atoi(argv[3]);
    std::vector<double> wiP((size_t)nc * ns * VLC_WINGPANEL_DOUBLES);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(wiP.data(), sizeof(double), wiP.size(), f) != wiP.size()) return 2;
    std::fclose(f);
    vlc::Rotor rotor(ctx, 0, 1, nc, ns, 2, 0);
    rotor.put_wing(0, wiP.data());
    const int N = rotor.N();
    std::vector<double> A((size_t)N * N), rhs(N), g(N);
    rotor.calcAIC(A.data());
    for (int i = 0; i < N; ++i) rhs[i] = i + 1.0;
    rotor.solve(rhs.data(), g.data());
    for (int r = 0; r < N; ++r) {
      for (int c = 0; c < N; ++c) std::printf("%.17g ", A[r + (size_t)N * c]);
      std::printf("\n");
    }
    for (int i = 0; i < N; ++i) std::printf("%.17g ", g[i]);
    std::printf("\n");
    return 0;
  } catch (const vlc::Error& e) {
    std::printf("error %d: %s\n", e.code, e.what());
    return 3;
  }
}
