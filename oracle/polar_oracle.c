/* CPU ORACLE in C (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's polar hot path, used (a) as the checker for the CUDA
 * kernels at batch sizes the numpy oracle cannot finish in seconds, and (b) as the timed CPU
 * baseline ("port") in bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it against the fixtures produced by
 * the unmodified reference (tests/golden/, oracle/gen_golden.py).  SC decisions are bit-exact; SCL
 * path metrics use libm exp/log, which may differ from numpy's by <= 1 ulp per term (best path and
 * its PM agree; see SURVEY 8c for why the tail of the list is ill-conditioned in the reference itself).
 *
 * Citations are reference file:line (/root/reference).
 * Threads: pthreads over disjoint codeword ranges (this image has no libgomp).
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC oracle/polar_oracle.c -o oracle/libpolar_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define LLR_MAX 30.0 /* polar_sc.py:21, polar_scl.py:35 */

static inline int ilog2(int n) { int m = 0; while ((1 << m) < n) ++m; return m; }

int oracle_num_threads(void) { long c = sysconf(_SC_NPROCESSORS_ONLN); return c > 0 ? (int)c : 1; }

typedef void (*range_fn)(void *ctx, long b0, long b1);
typedef struct { range_fn fn; void *ctx; long b0, b1; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->ctx, j->b0, j->b1); return 0; }
/* split [0,B) into `chunks` contiguous ranges handed out round-robin to nthreads workers */
static void run_parallel(range_fn fn, void *ctx, long B, int nthreads) {
  if (nthreads <= 0) nthreads = oracle_num_threads();
  if (nthreads > B) nthreads = (int)(B > 0 ? B : 1);
  if (nthreads <= 1) { fn(ctx, 0, B); return; }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].fn = fn; jobs[t].ctx = ctx;
    jobs[t].b0 = B * t / nthreads; jobs[t].b1 = B * (t + 1) / nthreads;
    pthread_create(&th[t], 0, job_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
  free(th); free(jobs);
}

/* ---------------------------------------------------------------- SC (fp32), polar_sc.py:33-112 */
static inline float clipf(float x) { return fminf(fmaxf(x, -(float)LLR_MAX), (float)LLR_MAX); }
static inline float sgnf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }
/* polar_sc.py:35-36,46: clip, then sign*sign*min|.| */
static inline float f32(float a, float b) {
  a = clipf(a); b = clipf(b);
  return sgnf(a) * sgnf(b) * fminf(fabsf(a), fabsf(b));
}
/* polar_sc.py:52: (1-2u)*x + y */
static inline float g32(float a, float b, uint8_t u) { return (1.0f - 2.0f * (float)u) * a + b; }

/* llr: stage buffers, stage s (node width 2^s) at llr + (2^s) .. ; beta: in-place partial sums */
static void sc_rec(float *stage_base, int s, int a, const uint8_t *frozen, uint8_t *u, uint8_t *beta) {
  float *L = stage_base + (1 << s); /* 2^s values of this node */
  if (s == 0) {
    uint8_t bit = 0;
    if (!frozen[a]) bit = (L[0] <= 0.0f) ? 1 : 0; /* polar_sc.py:90-98: frozen->0, llr==0 -> 1 */
    u[a] = bit; beta[a] = bit;
    return;
  }
  int h = 1 << (s - 1);
  float *C = stage_base + h; /* child buffer (stage s-1) */
  for (int j = 0; j < h; ++j) C[j] = f32(L[j], L[j + h]);          /* polar_sc.py:66-67 */
  sc_rec(stage_base, s - 1, a, frozen, u, beta);
  for (int j = 0; j < h; ++j) C[j] = g32(L[j], L[j + h], beta[a + j]); /* polar_sc.py:74-76 */
  sc_rec(stage_base, s - 1, a + h, frozen, u, beta);
  for (int j = 0; j < h; ++j) beta[a + j] ^= beta[a + h + j];        /* polar_sc.py:83-89 */
}

/* logit [B,n] fp32 (ln P1/P0); frozen [n] 0/1; u_out [B,n] all positions (frozen -> 0). */
typedef struct { const float *logit; const uint8_t *frozen; int n; uint8_t *u_out; } sc_ctx;
static void sc_range(void *vp, long b0, long b1) {
  sc_ctx *c = (sc_ctx *)vp; int n = c->n, m = ilog2(n);
  float *buf = (float *)malloc(sizeof(float) * 2 * (size_t)n);
  uint8_t *beta = (uint8_t *)malloc((size_t)n);
  for (long b = b0; b < b1; ++b) {
    for (int j = 0; j < n; ++j) buf[n + j] = -1.0f * c->logit[b * (long)n + j]; /* polar_sc.py:122 */
    sc_rec(buf, m, 0, c->frozen, c->u_out + b * (long)n, beta);
  }
  free(buf); free(beta);
}
void oracle_sc_decode(const float *logit, const uint8_t *frozen, int n, long B, uint8_t *u_out, int nthreads) {
  sc_ctx c = {logit, frozen, n, u_out};
  run_parallel(sc_range, &c, B, nthreads);
}

/* ------------------------------------------------------- SCL (fp64), polar_scl.py:49-209 */
static inline double clipd(double x) { return fmax(fmin(x, LLR_MAX), -LLR_MAX); } /* :96-97 */
static inline double sgnd(double x) { return (double)((x > 0.0) - (x < 0.0)); }
static inline double f64(double a, double b) { a = clipd(a); b = clipd(b); return sgnd(a) * sgnd(b) * fmin(fabs(a), fabs(b)); }
static inline double g64(double a, double b, uint8_t u) { return (1.0 - 2.0 * (double)u) * a + b; } /* :108 */
static inline double softplus_neg(double x) { return log(1.0 + exp(-x)); }                          /* :83 */

typedef struct {
  int n, m, L;
  const uint8_t *frozen;
  double *llr;   /* [L][2n]  per path stage buffers (same layout as SC)        */
  uint8_t *beta; /* [L][n]   per path in-place partial sums                    */
  uint8_t *u;    /* [L][n]   per path decisions                                */
  double *pm;    /* [L]                                                        */
  double *llr2; uint8_t *beta2; uint8_t *u2; double *pm2; /* scratch for forks */
} scl_t;

static void scl_fork(scl_t *S, int a) {
  int L = S->L, n = S->n;
  double cand[64]; int idx[64];
  for (int p = 0; p < L; ++p) {
    double x = clipd(S->llr[(size_t)p * 2 * n + 1]);           /* stage-0 value, polar_scl.py:81 */
    cand[p] = S->pm[p] + softplus_neg(x);                        /* u_hat=0 */
    cand[L + p] = S->pm[p] + softplus_neg(-x);                   /* u_hat=1 */
  }
  for (int i = 0; i < 2 * L; ++i) idx[i] = i;
  /* stable insertion sort ascending (reference: np.argsort default kind, tie order platform dependent) */
  for (int i = 1; i < 2 * L; ++i) {
    int v = idx[i]; int j = i - 1;
    while (j >= 0 && cand[idx[j]] > cand[v]) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = v;
  }
  for (int r = 0; r < L; ++r) {                                  /* polar_scl.py:109-120: full copy */
    int par = idx[r] % L; uint8_t bit = (uint8_t)(idx[r] / L);
    memcpy(S->llr2 + (size_t)r * 2 * n, S->llr + (size_t)par * 2 * n, sizeof(double) * 2 * n);
    memcpy(S->beta2 + (size_t)r * n, S->beta + (size_t)par * n, n);
    memcpy(S->u2 + (size_t)r * n, S->u + (size_t)par * n, n);
    S->u2[(size_t)r * n + a] = bit; S->beta2[(size_t)r * n + a] = bit;
    S->pm2[r] = cand[idx[r]];
  }
  double *t; uint8_t *tb;
  t = S->llr; S->llr = S->llr2; S->llr2 = t;
  tb = S->beta; S->beta = S->beta2; S->beta2 = tb;
  tb = S->u; S->u = S->u2; S->u2 = tb;
  t = S->pm; S->pm = S->pm2; S->pm2 = t;
}

static void scl_rec(scl_t *S, int s, int a) {
  int n = S->n, L = S->L;
  if (s == 0) {
    if (S->frozen[a]) {
      for (int p = 0; p < L; ++p) {
        double x = clipd(S->llr[(size_t)p * 2 * n + 1]);
        S->pm[p] += softplus_neg(x);                             /* polar_scl.py:82-83, u=0 */
        S->u[(size_t)p * n + a] = 0; S->beta[(size_t)p * n + a] = 0;
      }
    } else scl_fork(S, a);
    return;
  }
  int h = 1 << (s - 1);
  for (int p = 0; p < L; ++p) {
    double *Lp = S->llr + (size_t)p * 2 * n + (1 << s), *C = S->llr + (size_t)p * 2 * n + h;
    for (int j = 0; j < h; ++j) C[j] = f64(Lp[j], Lp[j + h]);   /* polar_scl.py:134-137 */
  }
  scl_rec(S, s - 1, a);
  for (int p = 0; p < L; ++p) {                                  /* S->llr may have been swapped */
    double *Lp = S->llr + (size_t)p * 2 * n + (1 << s), *C = S->llr + (size_t)p * 2 * n + h;
    uint8_t *bt = S->beta + (size_t)p * n;
    for (int j = 0; j < h; ++j) C[j] = g64(Lp[j], Lp[j + h], bt[a + j]); /* :140-144 */
  }
  scl_rec(S, s - 1, a + h);
  for (int p = 0; p < L; ++p) {
    uint8_t *bt = S->beta + (size_t)p * n;
    for (int j = 0; j < h; ++j) bt[a + j] ^= bt[a + h + j];    /* :147-153 */
  }
}

/* u_list [B,L,n] sorted by pm ascending; pm [B,L].  L <= 32. */
typedef struct { const float *logit; const uint8_t *frozen; int n, L; uint8_t *u_list; double *pm_out; } scl_ctx;
static void scl_range(void *vp, long b0, long b1) {
  scl_ctx *c = (scl_ctx *)vp;
  const float *logit = c->logit; const uint8_t *frozen = c->frozen; int n = c->n, L = c->L, m = ilog2(c->n);
  uint8_t *u_list = c->u_list; double *pm_out = c->pm_out;
  {
    scl_t S; S.n = n; S.m = m; S.L = L; S.frozen = frozen;
    S.llr = (double *)malloc(sizeof(double) * 2 * (size_t)n * L);
    S.llr2 = (double *)malloc(sizeof(double) * 2 * (size_t)n * L);
    S.beta = (uint8_t *)malloc((size_t)n * L); S.beta2 = (uint8_t *)malloc((size_t)n * L);
    S.u = (uint8_t *)malloc((size_t)n * L); S.u2 = (uint8_t *)malloc((size_t)n * L);
    S.pm = (double *)malloc(sizeof(double) * L); S.pm2 = (double *)malloc(sizeof(double) * L);
    for (long b = b0; b < b1; ++b) {
      for (int p = 0; p < L; ++p) {
        S.pm[p] = (p == 0) ? 0.0 : LLR_MAX;                      /* polar_scl.py:192-194 */
        double *top = S.llr + (size_t)p * 2 * n + n;
        for (int j = 0; j < n; ++j) top[j] = (double)(-1.0f * logit[b * (long)n + j]); /* :219, :200 */
        memset(S.u + (size_t)p * n, 0, n); memset(S.beta + (size_t)p * n, 0, n);
      }
      scl_rec(&S, m, 0);
      int idx[32];
      for (int i = 0; i < L; ++i) idx[i] = i;
      for (int i = 1; i < L; ++i) {                              /* final sort, polar_scl.py:204 */
        int v = idx[i]; int j = i - 1;
        while (j >= 0 && S.pm[idx[j]] > S.pm[v]) { idx[j + 1] = idx[j]; --j; }
        idx[j + 1] = v;
      }
      for (int r = 0; r < L; ++r) {
        memcpy(u_list + ((size_t)b * L + r) * n, S.u + (size_t)idx[r] * n, n);
        pm_out[b * (long)L + r] = S.pm[idx[r]];
      }
    }
    free(S.llr); free(S.llr2); free(S.beta); free(S.beta2); free(S.u); free(S.u2); free(S.pm); free(S.pm2);
  }
}
void oracle_scl_decode(const float *logit, const uint8_t *frozen, int n, int L, long B,
                       uint8_t *u_list, double *pm_out, int nthreads) {
  scl_ctx c = {logit, frozen, n, L, u_list, pm_out};
  run_parallel(scl_range, &c, B, nthreads);
}

/* ------------------------------------------------------- encoder, enc.py:33-42 */
void oracle_polar_transform(const uint8_t *u_full, int n, long B, uint8_t *c) {
  for (long b = 0; b < B; ++b) {
    uint8_t *x = c + b * (long)n;
    memcpy(x, u_full + b * (long)n, n);
    for (int s = 1; s < n; s <<= 1)
      for (int d = 0; d < n; ++d)
        if (!(d & s)) x[d] ^= x[d + s];
  }
}
