/*
 * flashpca_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic on FlashPCA2's blocked
 * partial-eigendecomposition hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (flashpca_b200/) never links or calls it.
 *
 * Parity pinning: the reference holds no golden vectors for this path
 * (SURVEY.md section 8c); its own tests compare against a dense
 * eigendecomposition computed at test time (flashpcaR/tests/testthat/
 * test_pca.R:45-105, HapMap3/test_pca.R:121-246).  This restatement is pinned
 * the same way: tests/test_oracle.py checks it against a dense numpy
 * eigh/XX' product on the reference's own bed fixtures and against the
 * session-probe constants recorded in SURVEY.md section 8c.  The reference
 * itself cannot be compiled here (Eigen, Spectra, Boost absent), so there is no
 * oracle/_ref binary.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the upstream flashpca tree).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FO_PACK_DENSITY 4
#define FO_PLINK_NA 3
#define FO_VAR_TOL 1e-9          /* util.h:33 */
#define FO_STANDARDISE_BINOM 2   /* util.h:36 */
#define FO_STANDARDISE_BINOM2 3  /* util.h:37 */

typedef struct {
   const unsigned char *bed;     /* packed genotypes, 3-byte header already skipped */
   unsigned long long N;         /* individuals */
   unsigned long long P;         /* nsnps */
   unsigned long long np;        /* bytes per SNP, ceil(N/4): data.cpp:169 */
   int stand_method;
   int use_preloaded;            /* data.cpp:293-297 */
   double *meansd;               /* P x 2 column-major: data.cpp:198 */
   double *lookup;               /* 4 x P column-major: data.cpp:200 */
   unsigned char *visited;       /* data.cpp:197 */
   unsigned char *tmp2;          /* unpack scratch, 4*np: data.cpp:194 */
   double *X;                    /* N x block_size column-major scratch */
   unsigned long long Xcols;
   double trace;                 /* svdwide.h:37 */
   int trace_done;
   unsigned int nops;
} fo_ctx;

/* data.cpp:65-126: byte -> 4 minor-allele dosages, 01 -> 3 (NA).
 * code = (byte >> 2q) & 3; dosage = !(code&1) + !(code>>1), except code==1. */
void fo_decode_plink(unsigned char *out, const unsigned char *in, unsigned int n)
{
   for(unsigned int i = 0; i < n; i++)
   {
      unsigned char b = in[i];
      for(int q = 0; q < 4; q++)
      {
	 unsigned char g = (b >> (2 * q)) & 3;
	 out[4 * i + q] = (g == 1) ? 3 : (unsigned char)(!(g & 1) + !(g >> 1));
      }
   }
}

/* data.cpp:128-148: byte -> 4 raw 2-bit codes. */
void fo_decode_plink_simple(unsigned char *out, const unsigned char *in, unsigned int n)
{
   for(unsigned int i = 0; i < n; i++)
      for(int q = 0; q < 4; q++)
	 out[4 * i + q] = (in[i] >> (2 * q)) & 3;
}

/* flashpca.cpp:636-686: --memory (MB) -> block_size; returns 0 on the
 * "memory not sufficient" paths. */
unsigned int fo_block_size_from_memory(unsigned long long N, unsigned long long P,
   unsigned int n_dim, int do_loadings, int memory_mb)
{
   long long mem = (long long)memory_mb * 1048576;
   long long req = 2 * (long long)P * 8 * 2 + 3 * (long long)P * 8
      + (long long)N * n_dim * 8 + (do_loadings ? (long long)P * n_dim * 8 : 0)
      + 2 * (long long)N + 2 * (long long)(N + P) * n_dim * 8
      + 2 * 1024 * 1024 + (long long)N * 8;
   long long remain = mem - req;
   if(remain <= 0)
      return 0;
   unsigned int bs = (unsigned int)floor(remain / ((double)N * 8.0));
   if(bs < 1)
      return 0;
   return bs < P ? bs : (unsigned int)P; /* flashpca.cpp:688 */
}

/* data.cpp:150-206 (get_size + prepare), serving the bed from memory. */
fo_ctx *fo_create(const unsigned char *bed_payload, unsigned long long N,
   unsigned long long P, int stand_method, const double *preloaded_meansd)
{
   fo_ctx *c = (fo_ctx *)calloc(1, sizeof(fo_ctx));
   c->bed = bed_payload;
   c->N = N;
   c->P = P;
   c->np = (N + FO_PACK_DENSITY - 1) / FO_PACK_DENSITY;
   c->stand_method = stand_method;
   c->meansd = (double *)calloc(2 * P, sizeof(double));
   c->lookup = (double *)calloc(4 * P, sizeof(double));
   c->visited = (unsigned char *)calloc(P, 1);
   c->tmp2 = (unsigned char *)malloc(c->np * FO_PACK_DENSITY);
   c->nops = 1;
   if(preloaded_meansd)
   {
      c->use_preloaded = 1;
      memcpy(c->meansd, preloaded_meansd, 2 * P * sizeof(double));
   }
   return c;
}

void fo_destroy(fo_ctx *c)
{
   if(!c)
      return;
   free(c->meansd);
   free(c->lookup);
   free(c->visited);
   free(c->tmp2);
   free(c->X);
   free(c);
}

/* data.cpp:257-322, one SNP's first-visit statistics + lookup row.
 * Returns -1 on an unknown standardisation method (data.cpp:283-288). */
static int fo_visit_snp(fo_ctx *c, unsigned long long k, unsigned char *scratch)
{
   const unsigned char *col = c->bed + c->np * k;
   double snp_avg = 0, sd;
   if(!c->use_preloaded)
   {
      unsigned int ngood = 0;
      fo_decode_plink(scratch, col, (unsigned int)c->np);
      for(unsigned long long i = 0; i < c->N; i++)
      {
	 if(scratch[i] != FO_PLINK_NA)
	 {
	    snp_avg += (double)scratch[i];
	    ngood++;
	 }
      }
      snp_avg /= ngood;
      double Pf = snp_avg / 2.0;
      if(c->stand_method == FO_STANDARDISE_BINOM)
	 sd = sqrt(Pf * (1 - Pf));
      else if(c->stand_method == FO_STANDARDISE_BINOM2)
	 sd = sqrt(2.0 * Pf * (1 - Pf));
      else
	 return -1;
      c->meansd[k] = snp_avg;
      c->meansd[c->P + k] = sd;
   }
   else
   {
      snp_avg = c->meansd[k];
      sd = c->meansd[c->P + k];
   }
   /* data.cpp:300-320: table is indexed by the RAW plink code. */
   if(sd > FO_VAR_TOL)
   {
      c->lookup[4 * k + 3] = (0 - snp_avg) / sd;
      c->lookup[4 * k + 2] = (1 - snp_avg) / sd;
      c->lookup[4 * k + 0] = (2 - snp_avg) / sd;
      c->lookup[4 * k + 1] = 0;
   }
   c->visited[k] = 1;
   return 0;
}

/* data.cpp:215-335: fill X (N x B, column-major) with standardised
 * genotypes of SNPs [start, stop] (inclusive). */
int fo_read_snp_block(fo_ctx *c, unsigned int start, unsigned int stop, double *X)
{
   unsigned int B = stop - start + 1;
   int rc = 0;
#pragma omp parallel
   {
      unsigned char *scratch = (unsigned char *)malloc(c->np * FO_PACK_DENSITY);
#pragma omp for schedule(static)
      for(unsigned int j = 0; j < B; j++)
      {
	 unsigned long long k = (unsigned long long)start + j;
	 if(!c->visited[k] && fo_visit_snp(c, k, scratch) != 0)
	    rc = -1;
	 fo_decode_plink_simple(scratch, c->bed + c->np * k, (unsigned int)c->np);
	 const double *lk = c->lookup + 4 * k;
	 double *xc = X + (unsigned long long)j * c->N;
	 for(unsigned long long i = 0; i < c->N; i++)
	    xc[i] = lk[scratch[i]];
      }
      free(scratch);
   }
   return rc;
}

static double *fo_block_buf(fo_ctx *c, unsigned int B)
{
   if(c->Xcols < B)
   {
      free(c->X);
      c->X = (double *)malloc(sizeof(double) * c->N * B);
      c->Xcols = B;
   }
   return c->X;
}

/* t = Xb' x : Xb is N x B column-major */
static void fo_gemv_t(const double *X, unsigned long long N, unsigned int B,
   const double *x, double *t)
{
#pragma omp parallel for schedule(static)
   for(unsigned int j = 0; j < B; j++)
   {
      const double *xc = X + (unsigned long long)j * N;
      double s = 0;
      for(unsigned long long i = 0; i < N; i++)
	 s += xc[i] * x[i];
      t[j] = s;
   }
}

/* y (+)= Xb t */
static void fo_gemv_n(const double *X, unsigned long long N, unsigned int B,
   const double *t, double *y, int accumulate)
{
   const unsigned long long CH = 2048;
   long long nchunk = (long long)((N + CH - 1) / CH);
#pragma omp parallel for schedule(static)
   for(long long ch = 0; ch < nchunk; ch++)
   {
      unsigned long long i0 = (unsigned long long)ch * CH;
      unsigned long long i1 = i0 + CH < N ? i0 + CH : N;
      if(!accumulate)
	 for(unsigned long long i = i0; i < i1; i++)
	    y[i] = 0;
      for(unsigned int j = 0; j < B; j++)
      {
	 const double *xc = X + (unsigned long long)j * N;
	 double tj = t[j];
	 for(unsigned long long i = i0; i < i1; i++)
	    y[i] += xc[i] * tj;
      }
   }
}

static double fo_sumsq(const double *X, unsigned long long n)
{
   double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
   for(long long i = 0; i < (long long)n; i++)
      s += X[i] * X[i];
   return s;
}

/* svdwide.cpp:21-68 (K == 1) and :71-118 / :229-275 (K > 1):
 * Y = sum_b X_b (X_b' Xin), blocks in order, first block assigned;
 * trace accumulated on the first call.  Xin, Y are N x K column-major. */
int fo_perform_op_multi(fo_ctx *c, const double *Xin, double *Y, unsigned int K,
   unsigned int block_size)
{
   unsigned long long P = c->P, N = c->N;
   if(block_size < 1 || block_size > P)
      block_size = (unsigned int)P;
   unsigned int nblocks = (unsigned int)((P + block_size - 1) / block_size); /* svdwide.h:57 */
   double *t = (double *)malloc(sizeof(double) * block_size);
   for(unsigned int b = 0; b < nblocks; b++)
   {
      unsigned int start = b * block_size;                 /* svdwide.h:63-68 */
      unsigned int stop = start + block_size - 1;
      if(stop >= P)
	 stop = (unsigned int)P - 1;
      unsigned int B = stop - start + 1;
      double *X = fo_block_buf(c, block_size);
      if(fo_read_snp_block(c, start, stop, X) != 0)
      {
	 free(t);
	 return -1;
      }
      for(unsigned int k = 0; k < K; k++)
      {
	 fo_gemv_t(X, N, B, Xin + (unsigned long long)k * N, t);
	 fo_gemv_n(X, N, B, t, Y + (unsigned long long)k * N, b > 0);
      }
      if(!c->trace_done)
      {
	 double s = fo_sumsq(X, N * B);
	 c->trace = (b == 0) ? s : c->trace + s;          /* svdwide.cpp:44-45,60-61 */
      }
   }
   c->trace_done = 1;
   c->nops++;
   free(t);
   return 0;
}

int fo_perform_op(fo_ctx *c, const double *x, double *y, unsigned int block_size)
{
   return fo_perform_op_multi(c, x, y, 1, block_size);
}

/* svdwide.cpp:122-153 / :157-188: Y (P x K) = X' Xin (N x K) */
int fo_crossprod_multi(fo_ctx *c, const double *Xin, double *Y, unsigned int K,
   unsigned int block_size)
{
   unsigned long long P = c->P, N = c->N;
   if(block_size < 1 || block_size > P)
      block_size = (unsigned int)P;
   unsigned int nblocks = (unsigned int)((P + block_size - 1) / block_size);
   for(unsigned int b = 0; b < nblocks; b++)
   {
      unsigned int start = b * block_size, stop = start + block_size - 1;
      if(stop >= P)
	 stop = (unsigned int)P - 1;
      unsigned int B = stop - start + 1;
      double *X = fo_block_buf(c, block_size);
      if(fo_read_snp_block(c, start, stop, X) != 0)
	 return -1;
      for(unsigned int k = 0; k < K; k++)
	 fo_gemv_t(X, N, B, Xin + (unsigned long long)k * N, Y + (unsigned long long)k * P + start);
   }
   c->nops++;
   return 0;
}

/* svdwide.cpp:193-226 / :312-343: Y (N x K) = X Vin (P x K) */
int fo_prod_multi(fo_ctx *c, const double *Vin, double *Y, unsigned int K,
   unsigned int block_size)
{
   unsigned long long P = c->P, N = c->N;
   if(block_size < 1 || block_size > P)
      block_size = (unsigned int)P;
   unsigned int nblocks = (unsigned int)((P + block_size - 1) / block_size);
   for(unsigned int b = 0; b < nblocks; b++)
   {
      unsigned int start = b * block_size, stop = start + block_size - 1;
      if(stop >= P)
	 stop = (unsigned int)P - 1;
      unsigned int B = stop - start + 1;
      double *X = fo_block_buf(c, block_size);
      if(fo_read_snp_block(c, start, stop, X) != 0)
	 return -1;
      for(unsigned int k = 0; k < K; k++)
	 fo_gemv_n(X, N, B, Vin + (unsigned long long)k * P + start,
	    Y + (unsigned long long)k * N, b > 0);
   }
   c->nops++;
   return 0;
}

double fo_get_trace(const fo_ctx *c) { return c->trace; }

void fo_get_meansd(const fo_ctx *c, double *out)
{
   memcpy(out, c->meansd, sizeof(double) * 2 * c->P);
}

void fo_get_lookup(const fo_ctx *c, double *out)
{
   memcpy(out, c->lookup, sizeof(double) * 4 * c->P);
}

int fo_num_threads(void)
{
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

void fo_set_num_threads(int n)
{
#ifdef _OPENMP
   omp_set_num_threads(n);
#else
   (void)n;
#endif
}

/* ---------------------------------------------------------------------------
 * Synthetic input for the CPU baseline legs of bench.py: the Balding-Nichols
 * generator of flashpca_b200/synth.py (counter-based hashing, SURVEY.md section
 * 8d), bit for bit, so the CPU arm is timed on SNP columns of the very matrix
 * the GPU arm holds.  Not part of the restated path; input preparation only.
 * thresholds: npop x nsnps_total uint32 row-major, pop: N bytes.
 * out: (j1 - j0) x ceil(N/4) packed PLINK bytes (no 3-byte header).
 * ------------------------------------------------------------------------- */
static unsigned long long fo_mix64(unsigned long long z)
{
   z += 0x9E3779B97F4A7C15ULL;
   z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
   z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
   return z ^ (z >> 31);
}

void fo_synth_bed(unsigned char *out, unsigned long long N, unsigned long long j0,
   unsigned long long j1, unsigned long long nsnps_total, const unsigned char *pop,
   const unsigned int *thresholds, unsigned int miss_thr, unsigned long long seed)
{
   const unsigned long long np = (N + 3) / 4;
   long long jj;
#pragma omp parallel for schedule(static)
   for(jj = (long long)j0; jj < (long long)j1; jj++)
   {
      const unsigned long long j = (unsigned long long)jj;
      unsigned char *row = out + (j - j0) * np;
      unsigned long long b;
      for(b = 0; b < np; b++)
      {
	 unsigned char byte = 0;
	 int q;
	 for(q = 0; q < 4; q++)
	 {
	    const unsigned long long i = 4 * b + q;
	    unsigned char code = 0;
	    if(i < N)
	    {
	       const unsigned long long h = fo_mix64(seed ^ fo_mix64(j * 0x100000001B3ULL + i));
	       const unsigned int u1 = (unsigned int)h, u2 = (unsigned int)(h >> 32);
	       const unsigned int um = (unsigned int)fo_mix64(h ^ 0xD6E8FEB86659FD93ULL);
	       const unsigned int t = thresholds[(unsigned long long)pop[i] * nsnps_total + j];
	       const int g = (u1 < t) + (u2 < t);
	       code = (g == 2) ? 0 : (g == 1 ? 2 : 3);
	       if(um < miss_thr)
		  code = 1;
	    }
	    byte |= (unsigned char)(code << (2 * q));
	 }
	 row[b] = byte;
      }
   }
}
