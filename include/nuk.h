/*
 * nuk.h -- C ABI of libnuk.so, the B200 (sm_100a) replacement for the native layer
 * (libbmas / SLEEF) underneath digikar99/numericals' data-parallel hot path.
 *
 * Two groups of entry points:
 *
 *  1. BMAS_<name>   (include/nuk_bmas.h, included below) -- the exact symbols the
 *     reference's translation tables bind through CFFI
 *     (src/basic-math/package.lisp:111-203, src/transcendental/package.lisp:62-84,
 *      src/linalg/package.lisp:39-43).  Same names, same argument meaning: element
 *     strides that may be 0 or negative, OUT may alias X or Y, reductions return by
 *     value.  Pointers may be DEVICE pointers (dense-arrays device storage: no copies)
 *     or plain HOST pointers (a pinned Lisp vector: the library stages
 *     host->HBM->host in pipelined chunks).  They run on the calling thread's current
 *     stream (nuk_set_stream).
 *
 *  2. nuk_*         -- new entry points that replace the reference's per-row Lisp
 *     loop nests (ptr-iterate-but-inner, src/utils/ptr-iterate-but-inner.lisp:3-71 and
 *     dense-numericals-src/utils/ptr-iterate-but-inner.lisp:3-86; with-simple-array-
 *     broadcast, src/utils/simple-array-broadcast.lisp:7-84) by ONE launch over a
 *     compact shape/stride descriptor, plus device storage, streams and events.
 *
 * All functions are re-entrant (the reference calls the C layer concurrently from
 * lparallel workers, src/transcendental/lparallel.lisp:70-96); per-thread state is
 * thread-local.  nuk_* functions return a status (0 = ok) and record a message
 * retrievable with nuk_last_error(); BMAS_* keep the reference's void/value returns
 * and record failures in the same place (nuk_last_status()).
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef NUK_H
#define NUK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUK_MAX_RANK 8

/* ------------------------------------------------------------------ status */
enum {
  NUK_OK = 0,
  NUK_ERR_INVALID = 1,     /* bad descriptor / argument */
  NUK_ERR_CUDA = 2,        /* CUDA runtime error (message holds cudaGetErrorString) */
  NUK_ERR_UNSUPPORTED = 3, /* (op, dtype) has no translation, like the reference's
                              "No CFUNCTION known for ..." (src/utils/translations.lisp:37) */
  NUK_ERR_NO_DEVICE = 4,
  NUK_ERR_COMM = 5         /* multi-GPU: no peer access, or a peer never arrived at a collective */
};
const char *nuk_last_error(void);
int nuk_last_status(void);
void nuk_clear_error(void);

/* ------------------------------------------------------------------ dtypes
 * Values = column index of the reference's translation table
 * (src/utils/translations.lisp:50-64, 66-82); 2 (cl) and 11 (fixnum) have no C type. */
enum {
  NUK_F32 = 0, NUK_F64 = 1,
  NUK_I64 = 3, NUK_I32 = 4, NUK_I16 = 5, NUK_I8 = 6,
  NUK_U64 = 7, NUK_U32 = 8, NUK_U16 = 9, NUK_U8 = 10
};
size_t nuk_dtype_size(int dtype); /* c-size, src/utils/translations.lisp:122-136 */

/* ------------------------------------------------------------------ ops */
enum { /* two-arg, same-type result: two-arg-fn/non-comparison, two-arg-fn/float */
  NUK_ADD = 0, NUK_SUB, NUK_MUL, NUK_DIV, NUK_MAX, NUK_MIN,
  NUK_POW, NUK_ATAN2,
  NUK_AND, NUK_OR, NUK_XOR, NUK_ANDNOT, NUK_SRA,
  /* log(x) / log(y) in ONE pass, bit-identical to the reference's three (nu:log x y = log y, log x, divide:
   * src/transcendental/two-arg-fn-float.lisp:531-547) -- SURVEY 8f-1; float dtypes only */
  NUK_LOGB,
  NUK_OP2_COUNT
};
enum { /* two-arg-fn/comparison -> (unsigned-byte 8) 0/1 */
  NUK_LT = 0, NUK_LE, NUK_EQ, NUK_NEQ, NUK_GE, NUK_GT, NUK_CMP_COUNT
};
enum { /* one-arg-fn/all, one-arg-fn/float, transcendental unaries */
  NUK_COPY = 0, NUK_ABS, NUK_ROUND, NUK_TRUNC, NUK_FLOOR, NUK_CEIL, NUK_SQRT,
  NUK_SIN, NUK_COS, NUK_TAN, NUK_ASIN, NUK_ACOS, NUK_ATAN,
  NUK_SINH, NUK_COSH, NUK_TANH, NUK_ASINH, NUK_ACOSH, NUK_ATANH,
  NUK_EXP, NUK_LOG, NUK_NOT,
  NUK_OP1_COUNT
};
enum { /* reductions: nu:sum nu:maximum nu:minimum (+ fused statistics) */
  NUK_SUM = 0, NUK_HMAX, NUK_HMIN,
  NUK_SUMSQ,  /* sum of x*x, each product rounded (no FMA): variance's second moment */
  NUK_RED_COUNT
};
/* flags for nuk_reduce */
enum {
  NUK_REDUCE_FAST = 0,      /* split the reduced axis, fixed combine order: deterministic */
  NUK_REDUCE_REF_ORDER = 1  /* strictly sequential along the axis: bit-identical to the
                               reference's row-accumulate order (src/basic-math/sum.lisp:69-71) */
};

/* ------------------------------------------------------------------ runtime */
int nuk_device_count(void);
int nuk_init(int device);            /* select device for the calling thread, warm the context */
int nuk_current_device(void);
const char *nuk_version(void);
int nuk_sm_count(void);

typedef void *nuk_stream_t;          /* a cudaStream_t; NULL = legacy default stream */
int nuk_stream_create(nuk_stream_t *s);
int nuk_stream_destroy(nuk_stream_t s);
int nuk_stream_sync(nuk_stream_t s);
int nuk_set_stream(nuk_stream_t s);  /* stream used by BMAS_* on this thread */
nuk_stream_t nuk_get_stream(void);
int nuk_device_sync(void);

typedef void *nuk_event_t;
int nuk_event_create(nuk_event_t *e);
int nuk_event_destroy(nuk_event_t e);
int nuk_event_record(nuk_event_t e, nuk_stream_t s);
int nuk_event_sync(nuk_event_t e);
int nuk_event_elapsed_ms(nuk_event_t start, nuk_event_t stop, float *ms);

/* device-resident storage for dense-arrays (replaces the Lisp heap vector of
 * standard-dense-array, dense-numericals-src/utils/primitives.lisp:3-57) */
int nuk_malloc(void **dptr, size_t bytes);
int nuk_free(void *dptr);
int nuk_memcpy_h2d(void *dst, const void *src, size_t bytes, nuk_stream_t s);
int nuk_memcpy_d2h(void *dst, const void *src, size_t bytes, nuk_stream_t s);
int nuk_memcpy_d2d(void *dst, const void *src, size_t bytes, nuk_stream_t s);
int nuk_memset(void *dst, int byte, size_t bytes, nuk_stream_t s);
int nuk_mem_info(size_t *free_bytes, size_t *total_bytes);
int nuk_host_alloc(void **hptr, size_t bytes);   /* pinned */
int nuk_host_free(void *hptr);
int nuk_host_register(void *hptr, size_t bytes); /* pin an existing buffer (e.g. a Lisp vector) */
int nuk_host_unregister(void *hptr);
int nuk_pointer_is_device(const void *p);        /* 1 device/managed, 0 host */
/* How BMAS_* symbols decide where an operand lives (thread-local).  NUK_PTR_AUTO asks the driver for
 * every pointer; NUK_PTR_DEVICE is for a host layer that keeps its arrays in device storage
 * (dense-arrays device backend): operands with a non-zero stride are device pointers by contract and
 * no driver query is made (a stride-0 operand may still be the reference's 1-element host scalar,
 * src/basic-math/two-arg-fn-non-comparison.lisp:226-232). */
enum { NUK_PTR_AUTO = 0, NUK_PTR_DEVICE = 1 };
int nuk_set_pointer_mode(int mode);
int nuk_get_pointer_mode(void);
/* number of host mover threads that stage pageable / strided host operands of BMAS_* calls
 * (option "stage_threads" / NUK_STAGE_THREADS, read when the pool starts; 0 = half the cores, at most 8) */
int nuk_host_mover_threads(void);
int nuk_flush_l2(nuk_stream_t s);                /* writes a > L2-sized scratch buffer (benchmark hygiene) */
/* diagnostic: counts float bit patterns in [lo_bits, hi_bits) for which the MUFU-seeded reciprocal /
   square root used inside the transcendental kernels differ from rcp.rn.f32 / sqrt.rn.f32 (expected 0
   over the normal range 2^-100 .. 2^100); synchronous */
int nuk_selftest_seeds(uint32_t lo_bits, uint32_t hi_bits, unsigned long long *mismatches, nuk_stream_t s);

/* CUDA-graph replay of a launch-bound call sequence (C1: 10^6-element maps).  Capture on a stream
 * made by nuk_stream_create after one warm-up pass of the same calls; maps and axis reductions are
 * capture-safe, value-returning BMAS reductions (they synchronise) are not. */
typedef void *nuk_graph_t;
int nuk_graph_begin(nuk_stream_t s);
int nuk_graph_end(nuk_stream_t s, nuk_graph_t *g);
int nuk_graph_launch(nuk_graph_t g, nuk_stream_t s);
int nuk_graph_destroy(nuk_graph_t g);

/* knobs (also readable from the environment at first use: NUK_GRID_MULT, NUK_STAGE_MB, ...).
 * NUK_LOG=1 in the environment prints one line per kernel launch to stderr; NUK_NVTX=1 wraps every entry point
 * in an NVTX range ("BMAS/map", "BMAS/reduce", "nuk/map2", "nuk/reduce", "nuk/comm_reduce" ...). */
int nuk_set_option(const char *name, long value);
long nuk_get_option(const char *name);
/* number of kernels launched by this library since load (all threads) */
unsigned long long nuk_launch_count(void);

/* ------------------------------------------------------------------ nd maps
 * dims[rank] with per-operand ELEMENT strides (0 on broadcast axes, negative allowed),
 * exactly what strides-for-broadcast (src/utils/broadcast.lisp:3-19) or a dense-arrays
 * view carries.  rank 0 = one element.  The library coalesces the descriptor (drops
 * size-1 axes, merges adjacent axes with compatible strides) and picks the contiguous,
 * row-broadcast or generic kernel.  x/y/out are DEVICE pointers to the element at
 * index (0,...,0); out may alias x and/or y exactly. */
int nuk_map1(int op, int dtype, int rank, const int64_t *dims, const int64_t *sx,
             const int64_t *so, const void *x, void *out, nuk_stream_t s);
int nuk_map2(int op, int dtype, int rank, const int64_t *dims, const int64_t *sx,
             const int64_t *sy, const int64_t *so, const void *x, const void *y, void *out,
             nuk_stream_t s);
int nuk_cmp2(int cmp, int dtype, int rank, const int64_t *dims, const int64_t *sx,
             const int64_t *sy, const int64_t *so, const void *x, const void *y,
             uint8_t *out, nuk_stream_t s);
int nuk_cast(int src_dtype, int dst_dtype, int rank, const int64_t *dims, const int64_t *sx,
             const int64_t *so, const void *x, void *out, nuk_stream_t s);
/* scalar operand passed by value (the reference writes it to a 1-element foreign
 * buffer and passes stride 0: src/basic-math/two-arg-fn-non-comparison.lisp:226-232).
 * scalar points to ONE host element of `dtype`; scalar_is_x selects the side. */
int nuk_map2_scalar(int op, int dtype, int rank, const int64_t *dims, const int64_t *sa,
                    const int64_t *so, const void *a, const void *host_scalar,
                    int scalar_is_x, void *out, nuk_stream_t s);
int nuk_cmp2_scalar(int cmp, int dtype, int rank, const int64_t *dims, const int64_t *sa,
                    const int64_t *so, const void *a, const void *host_scalar,
                    int scalar_is_x, uint8_t *out, nuk_stream_t s);
/* fill out with one host scalar (nu:fill / real -> array polymorphs) */
int nuk_fill(int dtype, int rank, const int64_t *dims, const int64_t *so,
             const void *host_scalar, void *out, nuk_stream_t s);

/* fused left fold out = (((x0 op x1) op x2) ...) over nin (2..8) contiguous same-type arrays of n
 * elements -- the n-ary nu:+ nu:- nu:* nu:/ nu:max nu:min in ONE pass instead of the reference's
 * nin-1 passes over OUT (src/basic-math/n-arg-fn.lisp:59-88); same fold order and roundings, so
 * bit-identical to the pass-by-pass fold.  out may alias any input. */
int nuk_fold(int op, int dtype, int64_t n, int nin, const void *const *xs, void *out, nuk_stream_t s);

/* ------------------------------------------------------------------ nd reductions
 * Reduce x over `axis` (or over everything when axis < 0) into out.
 * so[] are the element strides of OUT for the dims of x with `axis` removed (rank-1
 * entries; ignored for a full reduction, where out is one element).
 * Identities/sentinels follow the reference: sum -> type-zero, max -> type-min,
 * min -> type-max with FINITE float extremes (common.lisp:46-131). */
int nuk_reduce(int red, int dtype, int rank, const int64_t *dims, const int64_t *sx,
               int axis, const int64_t *so, const void *x, void *out, int flags,
               nuk_stream_t s);
/* nu:arg-maximum / nu:arg-minimum along `axis` (src/basic-math/arg-maximum.lisp:31-73): out is
 * int64, so[] its element strides for the dims of x with `axis` removed; index of the FIRST
 * extremum.  (The full, flattened form is BMAS_<t>himax / himin.) */
int nuk_argreduce(int is_max, int dtype, int rank, const int64_t *dims, const int64_t *sx, int axis,
                  const int64_t *so, const void *x, int64_t *out, nuk_stream_t s);
/* out[i] = out[i] / divisor, true IEEE division by a stride-0 scalar: the tail of
 * nu:mean (src/statistics/mean.lisp:90-97).  Float dtypes only. */
int nuk_div_scalar_inplace(int dtype, int64_t n, void *out, double divisor, nuk_stream_t s);
/* fused first/second raw moments in ONE pass over x: sum_out = sum(x), sumsq_out =
 * sum(x*x) with the same reduction tree as nuk_reduce. */
int nuk_moments(int dtype, int rank, const int64_t *dims, const int64_t *sx, int axis,
                const int64_t *so, const void *x, void *sum_out, void *sumsq_out, int flags,
                nuk_stream_t s);
/* variance tail, elementwise over n outputs, the reference's rounded steps
 * (src/statistics/variance.lisp:58-68): S=S*S; S=S/cnt; Q=Q-S; Q=Q/(cnt-ddof). Result in q. */
int nuk_variance_finish(int dtype, int64_t n, void *s_sum, void *q_sumsq, double count,
                        double ddof, nuk_stream_t s);

/* ------------------------------------------------------------------ multi-GPU (one box, <= 8 GPUs)
 * The path shards by contiguous slabs -- the reference's own thread-slicing rule lifted to devices:
 * ceiling(n / workers) units per worker, the last one takes what is left
 * (src/transcendental/lparallel.lisp:60-96), and for an nd array the FIRST axis that is long enough is the
 * one that is cut (dense-numericals-src/utils/lparallel.lisp:41-49).  Elementwise maps and reductions that
 * keep the cut axis are purely local (call the ordinary entry points on the slab).  What needs an exchange is
 * the combine step of a reduction over the cut axis; the last kernel of the reduction does it itself over
 * NVLink peer memory and folds the per-rank values IN RANK ORDER, so all ranks hold bit-identical results and
 * the result is reproducible for a given world size.
 *
 * One process, all GPUs (a Common Lisp image):  nuk_comm_init_all(n, NULL, comms); then per collective one
 * call PER communicator, all issued before any of them is waited for (they are asynchronous on the stream
 * given; each communicator's call switches to its device and back).
 * One process per GPU:  nuk_comm_create on the rank's device, nuk_comm_export -> exchange the
 * NUK_COMM_HANDLE_BYTES-byte handles with whatever the host has (rank order) -> nuk_comm_connect.
 * Every rank must issue the same sequence of collectives.  A peer that does not arrive within
 * NUK_COMM_TIMEOUT_MS (environment, default 20000) makes the kernels give up; nuk_comm_status and every
 * later collective then return NUK_ERR_COMM. */
typedef struct nuk_comm *nuk_comm_t;
#define NUK_COMM_HANDLE_BYTES 128
int nuk_shard_bounds(int64_t n, int world, int rank, int64_t *lo, int64_t *hi);
int nuk_shard_axis(int rank, const int64_t *dims, int world);   /* the axis to cut, -1 if none is long enough */
int nuk_comm_init_all(int ndev, const int *devices /* NULL = 0..ndev-1 */, nuk_comm_t *comms);
int nuk_comm_create(int world, int rank, nuk_comm_t *comm);      /* on the calling thread's device */
int nuk_comm_export(nuk_comm_t comm, void *handle_out);
int nuk_comm_connect(nuk_comm_t comm, const void *all_handles);  /* world * NUK_COMM_HANDLE_BYTES, rank order */
int nuk_comm_destroy(nuk_comm_t comm);
int nuk_comm_world(nuk_comm_t comm);
int nuk_comm_rank(nuk_comm_t comm);
int nuk_comm_device(nuk_comm_t comm);
int nuk_comm_status(nuk_comm_t comm);
/* out[i] = local_0[i] (+) local_1[i] (+) ... (+) local_{W-1}[i] for i < count, (+) = sum / max / min of `red`;
 * local and out are device arrays of this rank (out may be local).  The combine of an axis reduction over the
 * cut axis: nuk_reduce on the slab into `local`, then this. */
int nuk_comm_allreduce(nuk_comm_t comm, int red, int dtype, int64_t count, const void *local, void *out,
                       nuk_stream_t s);
/* FULL reduction of the whole sharded array: x describes this rank's slab (dims / element strides as in
 * nuk_reduce); out receives the global result on every rank (one device element).  Two launches: pass 1 over
 * the slab, then pass 2 + exchange + rank-order fold in one kernel. */
int nuk_comm_reduce(nuk_comm_t comm, int red, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                    const void *x, void *out, nuk_stream_t s);
/* global sum(x) and sum(x*x) in one pass over the slab (nu:variance / nu:std on a sharded array) */
int nuk_comm_moments(nuk_comm_t comm, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                     const void *x, void *sum_out, void *sumsq_out, nuk_stream_t s);

/* ------------------------------------------------------------------ resolver
 * The "utils broadcast/stride resolver": numpy right-aligned rule of
 * %broadcast-compatible-p / strides-for-broadcast (src/utils/broadcast.lisp:3-63).
 * Inputs: nargs shapes (ranks[i], dims_i, strides_i in elements; strides_i may be NULL
 * = row-major).  Outputs: broadcast rank/dims and per-argument broadcast strides
 * (out_strides is nargs*NUK_MAX_RANK, row i at out_strides + i*NUK_MAX_RANK).
 * Returns NUK_ERR_INVALID for incompatible-broadcast-dimensions. */
int nuk_broadcast_resolve(int nargs, const int *ranks, const int64_t *const *dims,
                          const int64_t *const *strides, int *out_rank, int64_t *out_dims,
                          int64_t *out_strides);
/* descriptor compaction used by every nd entry point (exposed for tests): drops size-1
 * axes and merges adjacent axes when every operand allows it. strides is
 * nops*NUK_MAX_RANK in/out. Returns the new rank. */
int nuk_coalesce(int nops, int rank, int64_t *dims, int64_t *strides);

#ifdef __cplusplus
}
#endif

#include "nuk_bmas.h"

#endif /* NUK_H */
