/*
 * oracle/dgemm_oracle.c -- CPU restatement of EigenCuda's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (eigencuda_b200/) never links, imports or calls it.
 *
 * What it restates (citations are into /root/reference):
 *   - the arithmetic of CudaPipeline::gemm, src/cudapipeline.cc:17-32: one FP64
 *     C(m x n) = 1.0 * A(m x k) * B(k x n) + 0.0 * C, all column-major, lda = A.rows,
 *     ldb = B.rows, ldc = C.rows (the cublasDgemm argument pattern at :29-31);
 *   - the shape check that precedes it, src/cudapipeline.cc:26-28 (inner dimensions only);
 *   - the streaming loop the reference is driven in, README.md:62-66 == src/tests/test_dot.cc:81-85:
 *     for each tensor element T_i: copy_to_gpu(T_i); gemm(T_i, A, C); results[i] = C;
 *   - the comparison the reference's tests use, Eigen isApprox (src/tests/test_dot.cc:31,50,87-89):
 *     ||a-b||_F^2 <= p^2 * min(||a||_F^2, ||b||_F^2), p = 1e-12 for double.
 * The arithmetic itself lives in closed-source cuBLAS (cublasDgemm, version not pinned by the
 * reference; this image ships 12.9.1.4), so what is restated is the published BLAS-3 DGEMM
 * definition, generalised with op(A)/op(B)/alpha/beta exactly as cublasDgemm defines them.
 *
 * Pinning: checked against the reference's own known-answer vectors (A,B,C,D -> X,Y,Z of
 * src/tests/test_dot.cc:53-90) in tests/test_oracle.py, and against the unmodified reference
 * sources compiled into oracle/_ref (cuBLAS on the GPU box) in tests/test_gpu_parity.py.
 *
 * Three accumulation flavours, identical semantics:
 *   oracle_dgemm       plain double, k-ordered sum           (what "a CPU product" means)
 *   oracle_dgemm_ld    long double (x87, 64-bit mantissa)    (accuracy truth, ~1e-19)
 *   oracle_dgemm_fast  cache-blocked + OpenMP, double        (speed baseline, kind = "port")
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_OK 0
#define ORACLE_SHAPE_MISMATCH 1
#define ORACLE_BAD_ARG 2

/* op(X)(i,j) for a column-major X with leading dimension ld. */
static inline double at(const double *x, int64_t ld, int trans, int64_t i, int64_t j) {
  return trans ? x[i * ld + j] : x[j * ld + i];
}

/* src/cudapipeline.cc:26-28: only the inner dimensions are validated. */
int oracle_gemm_shape_check(int64_t a_rows, int64_t a_cols, int64_t b_rows, int64_t b_cols) {
  (void)a_rows;
  (void)b_cols;
  return a_cols != b_rows ? ORACLE_SHAPE_MISMATCH : ORACLE_OK;
}

int oracle_dgemm(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                 const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *C,
                 int64_t ldc) {
  if (m < 0 || n < 0 || k < 0) return ORACLE_BAD_ARG;
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < m; ++i) {
      double acc = 0.0;
      for (int64_t p = 0; p < k; ++p) acc += at(A, lda, transa, i, p) * at(B, ldb, transb, p, j);
      /* beta == 0 must not read C (cuBLAS semantics: NaNs in C are not propagated). */
      C[j * ldc + i] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * C[j * ldc + i];
    }
  return ORACLE_OK;
}

int oracle_dgemm_ld(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                    const double *A, int64_t lda, const double *B, int64_t ldb, double beta,
                    double *C, int64_t ldc) {
  if (m < 0 || n < 0 || k < 0) return ORACLE_BAD_ARG;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < m; ++i) {
      long double acc = 0.0L;
      for (int64_t p = 0; p < k; ++p)
        acc += (long double)at(A, lda, transa, i, p) * (long double)at(B, ldb, transb, p, j);
      long double r = (long double)alpha * acc;
      if (beta != 0.0) r += (long double)beta * (long double)C[j * ldc + i];
      C[j * ldc + i] = (double)r;
    }
  return ORACLE_OK;
}

/* Cache-blocked, OpenMP over column panels.  Same result definition; summation order differs
 * from oracle_dgemm (k is blocked), as it does in any optimised BLAS. */
int oracle_dgemm_fast(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                      const double *A, int64_t lda, const double *B, int64_t ldb, double beta,
                      double *C, int64_t ldc) {
  if (m < 0 || n < 0 || k < 0) return ORACLE_BAD_ARG;
  enum { MB = 96, NB = 8, KB = 256 };
#pragma omp parallel
  {
    double *apack = (double *)malloc(sizeof(double) * MB * KB);
#pragma omp for schedule(dynamic, 4)
    for (int64_t j0 = 0; j0 < n; j0 += NB) {
      int64_t nb = n - j0 < NB ? n - j0 : NB;
      for (int64_t j = 0; j < nb; ++j)
        for (int64_t i = 0; i < m; ++i)
          C[(j0 + j) * ldc + i] = (beta == 0.0) ? 0.0 : beta * C[(j0 + j) * ldc + i];
      for (int64_t p0 = 0; p0 < k; p0 += KB) {
        int64_t kb = k - p0 < KB ? k - p0 : KB;
        for (int64_t i0 = 0; i0 < m; i0 += MB) {
          int64_t mb = m - i0 < MB ? m - i0 : MB;
          /* pack op(A)[i0:i0+mb, p0:p0+kb] as [p][i], unit stride in i */
          for (int64_t p = 0; p < kb; ++p)
            for (int64_t i = 0; i < mb; ++i) apack[p * MB + i] = at(A, lda, transa, i0 + i, p0 + p);
          for (int64_t j = 0; j < nb; ++j) {
            double acc[MB];
            for (int64_t i = 0; i < mb; ++i) acc[i] = 0.0;
            for (int64_t p = 0; p < kb; ++p) {
              double b = alpha * at(B, ldb, transb, p0 + p, j0 + j);
              const double *ap = apack + p * MB;
              for (int64_t i = 0; i < mb; ++i) acc[i] += ap[i] * b;
            }
            double *c = C + (j0 + j) * ldc + i0;
            for (int64_t i = 0; i < mb; ++i) c[i] += acc[i];
          }
        }
      }
    }
    free(apack);
  }
  return ORACLE_OK;
}

/* ---- deterministic synthetic inputs ----------------------------------
 * x = 2u - 1, u uniform in [0,1) with 53 random bits, from splitmix64 keyed on
 * (seed, matrix_id, linear column-major index).  The CUDA generator in
 * eigencuda_b200/csrc/ uses the same integer recipe, so host and device data are bit-identical.
 * Distribution matches the reference tests' MatrixXd::Random (uniform [-1,1]),
 * src/tests/test_dot.cc:37-38. */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

double oracle_uniform_at(uint64_t seed, uint64_t matrix_id, uint64_t index) {
  uint64_t key = splitmix64(seed ^ splitmix64(matrix_id + 0x632BE59BD9B4E019ULL));
  uint64_t r = splitmix64(key + index * 0xD1342543DE82EF95ULL);
  double u = (double)(r >> 11) * (1.0 / 9007199254740992.0); /* 2^-53 */
  return 2.0 * u - 1.0;
}

void oracle_fill_uniform(uint64_t seed, uint64_t matrix_id, double *dst, int64_t count) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < count; ++i) dst[i] = oracle_uniform_at(seed, matrix_id, (uint64_t)i);
}

/* `count` generator values at indices start, start+stride, ...: one row or one column of a large
 * synthetic matrix without materialising it (sampled verification of the n = 65536 product). */
void oracle_fill_uniform_strided(uint64_t seed, uint64_t matrix_id, uint64_t start, uint64_t stride,
                                 double *dst, int64_t count) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < count; ++i)
    dst[i] = oracle_uniform_at(seed, matrix_id, start + (uint64_t)i * stride);
}

/* ---- comparison ------------------------------------------------------------------------------- */
/* ||a-b||_F / min(||a||_F, ||b||_F): Eigen's isApprox(a,b,p) is exactly (this <= p), p = 1e-12
 * by default for double (src/tests/test_dot.cc:31,50,87-89 use the default). */
double oracle_rel_frobenius(const double *a, const double *b, int64_t count) {
  long double d = 0, na = 0, nb = 0;
  for (int64_t i = 0; i < count; ++i) {
    long double x = a[i], y = b[i];
    d += (x - y) * (x - y);
    na += x * x;
    nb += y * y;
  }
  long double mn = na < nb ? na : nb;
  if (mn == 0) return d == 0 ? 0.0 : INFINITY;
  return (double)sqrtl(d / mn);
}

int oracle_is_approx(const double *a, const double *b, int64_t count, double prec) {
  return oracle_rel_frobenius(a, b, count) <= prec;
}

/* ---- the streaming loop (README.md:62-66, src/tests/test_dot.cc:81-85) --------------------------
 * mode 0: results[i] = T_i * F      (the reference's order: gemm(cuma_B, cuma_A, cuma_C))
 * mode 1: results[i] = F * T_i
 * mode 2: results[i] = F^T * T_i
 * mode 3: results[i] = F^T * T_i * F (basis transform; two gemms, BASELINE.json config 4)
 * T_i is t_rows x t_cols, F is f_rows x f_cols, all column-major with ld = rows.
 * flavour 0 = plain double, 1 = long double, 2 = blocked/OpenMP.                                  */
typedef int (*gemm_fn)(int, int, int64_t, int64_t, int64_t, double, const double *, int64_t,
                       const double *, int64_t, double, double *, int64_t);

int oracle_tensor_stream(int mode, int flavour, const double *F, int64_t f_rows, int64_t f_cols,
                         const double *const *tensor, double *const *results, int64_t count,
                         int64_t t_rows, int64_t t_cols) {
  gemm_fn g = flavour == 1 ? oracle_dgemm_ld : flavour == 2 ? oracle_dgemm_fast : oracle_dgemm;
  double *tmp = NULL;
  int rc = ORACLE_OK;
  switch (mode) {
    case 0: rc = oracle_gemm_shape_check(t_rows, t_cols, f_rows, f_cols); break;
    case 1: rc = oracle_gemm_shape_check(f_rows, f_cols, t_rows, t_cols); break;
    case 2: rc = oracle_gemm_shape_check(f_cols, f_rows, t_rows, t_cols); break;
    case 3:
      rc = oracle_gemm_shape_check(f_cols, f_rows, t_rows, t_cols);
      if (!rc) rc = oracle_gemm_shape_check(f_cols, t_cols, f_rows, f_cols);
      if (!rc) tmp = (double *)malloc(sizeof(double) * (size_t)(f_cols * t_cols + 1));
      break;
    default: return ORACLE_BAD_ARG;
  }
  if (rc) return rc;
  for (int64_t i = 0; i < count && !rc; ++i) {
    switch (mode) {
      case 0: rc = g(0, 0, t_rows, f_cols, t_cols, 1.0, tensor[i], t_rows, F, f_rows, 0.0, results[i], t_rows); break;
      case 1: rc = g(0, 0, f_rows, t_cols, f_cols, 1.0, F, f_rows, tensor[i], t_rows, 0.0, results[i], f_rows); break;
      case 2: rc = g(1, 0, f_cols, t_cols, f_rows, 1.0, F, f_rows, tensor[i], t_rows, 0.0, results[i], f_cols); break;
      case 3:
        rc = g(1, 0, f_cols, t_cols, f_rows, 1.0, F, f_rows, tensor[i], t_rows, 0.0, tmp, f_cols);
        if (!rc) rc = g(0, 0, f_cols, f_cols, t_cols, 1.0, tmp, f_cols, F, f_rows, 0.0, results[i], f_cols);
        break;
    }
  }
  free(tmp);
  return rc;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
