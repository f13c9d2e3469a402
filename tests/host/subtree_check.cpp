// Host-side check of the register-subtree SC decoder (polar_common.cuh, compiled as plain C++)
// against the C oracle.  Build/run by tests/test_host_subtree.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "polar_common.cuh"
extern "C" void oracle_sc_decode(const float*, const uint8_t*, int, long, uint8_t*, int);

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }

template <int T> int check(int iters) {
  constexpr int N = 1 << T;
  int bad = 0;
  for (int it = 0; it < iters; ++it) {
    float x[N]; uint8_t fz[N]; uint32_t fm = 0;
    int mode = it % 6;
    for (int j = 0; j < N; ++j) {
      float v = ((int)(rnd() % 2001) - 1000) / 100.0f;           // [-10,10]
      if (mode == 1) v = roundf(v);                               // quantised: ties and zeros
      if (mode == 2) v *= 10.0f;                                  // exercises the +-30 clip
      if (mode == 3 && rnd() % 4 == 0) v = 0.0f;
      x[j] = v;
      int f;
      if (mode == 4) f = 0; else if (mode == 5) f = (j < N / 2); else f = rnd() & 1;
      if (it % 11 == 0) f = 1;
      fz[j] = (uint8_t)f; fm |= (uint32_t)f << j;
    }
    uint32_t u = 0; polar::SubTree<T>::run(x, fm, u);
    float logit[N]; for (int j = 0; j < N; ++j) logit[j] = -x[j];  // oracle negates logits
    uint8_t uo[N]; oracle_sc_decode(logit, fz, N, 1, uo, 1);
    uint32_t ur = 0; for (int j = 0; j < N; ++j) ur |= (uint32_t)uo[j] << j;
    if (u != ur) { if (bad < 5) printf("T=%d it=%d mode=%d mismatch %08x vs %08x\n", T, it, mode, u, ur); ++bad; }
  }
  return bad;
}
// BetaTree<5> (the partial-sum-only recursion the kernels run): equal to SubTree's partial sums; with cl = true on inputs
// inside [-30, 30] (outputs of an f) the clip-free first level must not change anything
static int check_beta(int iters) {
  int bad = 0;
  for (int it = 0; it < iters; ++it) {
    float x[32]; uint32_t fm = 0;
    const int mode = it % 5;
    for (int j = 0; j < 32; ++j) {
      float v = ((int)(rnd() % 8001) - 4000) / 100.0f;           // [-40, 40]: g outputs of the levels below exceed the clip
      if (mode == 1) v = roundf(v);
      if (mode == 2 && rnd() % 4 == 0) v = 0.0f;
      x[j] = v;
      int f = (mode == 3) ? 0 : (mode == 4 ? (j < 16) : (int)(rnd() & 1));
      fm |= (uint32_t)f << j;
    }
    uint32_t u = 0;
    const uint32_t want = polar::SubTree<5>::run(x, fm, u);
    if (polar::BetaTree<5>::run(x, fm) != want) { if (bad < 5) printf("BetaTree it=%d mismatch\n", it); ++bad; }
    float xc[32];
    for (int j = 0; j < 32; ++j) xc[j] = fminf(fmaxf(x[j], -30.0f), 30.0f);
    const uint32_t wc = polar::SubTree<5>::run(xc, fm, u);
    if (polar::BetaTree<5, true>::run(xc, fm) != wc || polar::BetaTree<5>::run(xc, fm) != wc) {
      if (bad < 5) printf("BetaTree cl it=%d mismatch\n", it);
      ++bad;
    }
  }
  return bad;
}
int main() {
  int bad = 0;
  bad += check_beta(40000);
  bad += check<1>(2000); bad += check<2>(4000); bad += check<3>(8000); bad += check<4>(8000); bad += check<5>(20000);
  // transform involution
  for (int i = 0; i < 1000; ++i) { uint32_t v = rnd(); if (polar::ptransform<5>(polar::ptransform<5>(v)) != v) ++bad; }
  printf("subtree_check bad=%d\n", bad);
  return bad != 0;
}
