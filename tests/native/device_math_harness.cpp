// Host harness for scalar device math that needs no GPU to be checked (tests/test_device_math_host.py pastes the
// functions, cut verbatim out of csrc/attn_fwd.cu and csrc/gemm.cu, at the markers):
//   ex2_poly        2^x on the FMA pipe (attention forward: one pair of every eight exponentials)
//   gelu_tanh(_grad) the FF activation and its derivative as the GEMM epilogues apply them (tanh.approx -> tanhf here)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define __device__
#define __forceinline__ inline
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float tanh_fast(float x) { return std::tanh(x); }

/*@@EX2_POLY@@*/
/*@@GELU@@*/

int main() {
  // 2^x for x in [-120, 0]: relative error against exp2 in double
  double worst = 0.0, worst_x = 0.0;
  for (int i = 0; i <= 12000000; ++i) {
    const float x = -(float)i * 1e-5f;
    const double want = std::exp2((double)x);
    const double rel = std::fabs((double)ex2_poly(x) - want) / want;
    if (rel > worst) { worst = rel; worst_x = x; }
  }
  std::printf("ex2_poly max_rel %.3e at %.5f\n", worst, worst_x);
  // below the clamp the result stays tiny and finite (P of a masked key), and 2^0 is exactly 1
  std::printf("ex2_poly clamp %d one %d\n", (int)(ex2_poly(-1e30f) > 0.f && ex2_poly(-1e30f) < 1e-35f), (int)(ex2_poly(0.f) == 1.0f));
  // GELU' against a central difference of GELU in double
  double gworst = 0.0;
  for (int i = -80000; i <= 80000; ++i) {
    const float x = (float)i * 1e-4f;
    auto g = [](double v) { return 0.5 * v * (1.0 + std::tanh(0.7978845608028654 * (v + 0.044715 * v * v * v))); };
    const double want = (g((double)x + 1e-6) - g((double)x - 1e-6)) / 2e-6;
    gworst = std::fmax(gworst, std::fabs((double)gelu_tanh_grad(x) - want));
    gworst = std::fmax(gworst, std::fabs((double)gelu_tanh(x) - g((double)x)));
  }
  std::printf("gelu max_abs %.3e\n", gworst);
  return 0;
}
