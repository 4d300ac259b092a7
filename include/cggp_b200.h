/* cggp_b200.h - C ABI of the B200-native conjugate-gradient hot path (libcggp_b200.so).
 *
 * The reference (awav/conjugate-gradient-sparse-gp) is pure Python: it has no FFI, plugin or operator registry.
 * Its boundary for this path is three Python surfaces, each of which one group of entry points
 * below replaces; the Python mirror in cggp_b200/ keeps the reference names/argument order and calls these
 * through ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer marked `dev` is a DEVICE pointer borrowed from the caller (e.g. a DLPack capsule); the
 *     library never frees or retains it beyond the call unless stated.  `host` pointers are host memory.
 *   - matrices are row-major with explicit leading dimensions (in elements).
 *   - dtype is CGGP_F64 or CGGP_F32 and applies to every floating-point buffer of the call.
 *   - every call is asynchronous on the ctx stream unless stated, returns 0 on success and <0 on error;
 *     cggp_last_error(ctx) gives the message.  A ctx is bound to one device and one stream and is not re-entrant.
 *   - there is no CPU fallback: without a CUDA device cggp_ctx_create fails.
 */
#ifndef CGGP_B200_H
#define CGGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cggp_ctx cggp_ctx;

enum cggp_dtype { CGGP_F32 = 0, CGGP_F64 = 1 };

/* GPflow stationary kernels the reference instantiates (cggp/cli_utils.py:103-135,363-368,462-473). */
enum cggp_kernel_kind { CGGP_SE = 0, CGGP_MATERN12 = 1, CGGP_MATERN32 = 2, CGGP_MATERN52 = 3 };

/* cggp/distance.py:9-34 distance types (+ GPflow square_distance as used by cggp/optimize.py:50). */
enum cggp_distance { CGGP_DIST_EUCLIDEAN = 0, CGGP_DIST_COVARIANCE = 1, CGGP_DIST_CORRELATION = 2,
                     CGGP_DIST_SQEUCLIDEAN = 3 };

/* Output selector of cggp_kernel_matrix. */
enum cggp_tile_output { CGGP_OUT_KERNEL = 0, CGGP_OUT_DISTANCE = 1 };

enum cggp_status {
  CGGP_OK = 0, CGGP_ERR_INVALID = -1, CGGP_ERR_CUDA = -2, CGGP_ERR_UNSUPPORTED = -3, CGGP_ERR_COMM = -4
};

/* ---------------------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------------------- */
int cggp_ctx_create(int device, cggp_ctx** out);
int cggp_ctx_destroy(cggp_ctx* ctx);
/* Bind the CUDA stream (cudaStream_t as void*) that all later calls launch on; default is the NULL stream. */
int cggp_ctx_set_stream(cggp_ctx* ctx, void* cuda_stream);
const char* cggp_last_error(cggp_ctx* ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
int64_t cggp_launch_count(cggp_ctx* ctx);
const char* cggp_version(void);

/* Per-section device timing for bench.py's roofline (CUDA events on the ctx stream around every launch group):
 * section 0 = matrix-free Kuf Kfu product, 1 = dense symmetric product (Kuu / A), 2 = fused CG vector step,
 * 3 = all-reduce.  enable(on != 0) clears the accumulated records; read() synchronises the stream and returns the
 * summed milliseconds and the number of timed launch groups of that section. */
#define CGGP_PROF_SECTIONS 4
int cggp_profile_enable(cggp_ctx* ctx, int on);
int cggp_profile_read(cggp_ctx* ctx, int section, double* host_ms_total, int64_t* host_count);

/* Multi-GPU: one rank per GPU; the only collective on the path is the per-iteration all-reduce of the
 * partial [B, M] product.  The id is an ncclUniqueId (128 bytes) made on rank 0 and sent to the others by the
 * host (torch.distributed).  NCCL is dlopen'ed from the process (the torch-bundled libnccl.so.2). */
int cggp_comm_unique_id(void* host_id128);
int cggp_ctx_comm_init(cggp_ctx* ctx, const void* host_id128, int rank, int world);
int cggp_ctx_comm_destroy(cggp_ctx* ctx);
/* In-place sum all-reduce of `count` elements on the ctx stream (no-op when world == 1): ncclAllReduce, or - when
 * the peer buffers below have been set up and the vector fits a slot - ONE kernel over NVLink peer memory: every rank
 * publishes its vector in its own IPC-shared buffer and sums all ranks' slots in rank order (bit-identical on all
 * ranks).  Measured on 2 B200 for the 32 KiB vector of c3: 31 us against NCCL's 24 us per call, both dominated by
 * waiting for the slower rank; the Python layer therefore sets the peers up only on request (CGGP_PEER_ALLREDUCE=1). */
int cggp_allreduce_sum(cggp_ctx* ctx, int dtype, void* dev_buf, int64_t count);
/* Peer buffers of the one-shot all-reduce (after cggp_ctx_comm_init; ranks of ONE node): cggp_peer_alloc allocates
 * this rank's buffer (two slots of slot_bytes + flags) and writes its 64-byte CUDA IPC handle; the host sends the
 * handles round (torch.distributed all_gather); cggp_peer_open maps all ranks' buffers (world x 64 bytes, rank order). */
int cggp_peer_alloc(cggp_ctx* ctx, int64_t slot_bytes, void* host_handle64);
int cggp_peer_open(cggp_ctx* ctx, const void* host_handles, int world);
int cggp_peer_close(cggp_ctx* ctx);
int cggp_peer_enabled(cggp_ctx* ctx);

/* ---------------------------------------------------------------------------------------------------------
 * Kernel evaluation  (replaces the GPflow calls at cggp/models.py:112,141-143,236,255-257,300,333-335 and
 * cggp/distance.py:17-20,26-29: Stationary.scale, square_distance, K_r2 / K_r, Kuu, Kuf)
 * ------------------------------------------------------------------------------------------------------- */

/* "Prepared points": P[i, 0:D] = X[i, :] / lengthscales, P[i, D] = 1 (the spare column: it carries the |z|^2 term
 * through the fused matvec's DMMA), zero padding up to ldp (ldp = D + 1 rounded up to a multiple of 4, see
 * cggp_prepared_ld), norms[i] = sum_d P[i, d]^2.
 * lengthscales: host double[D] (ARD) or host double[1] with ls_count = 1 (isotropic). */
int64_t cggp_prepared_ld(int D);
int cggp_prepare_points(cggp_ctx* ctx, int dtype, const void* dev_X, int64_t n, int D, int64_t ldx,
                        const double* host_lengthscales, int ls_count, void* dev_P, int64_t ldp,
                        void* dev_norms);

/* out[i, j] for prepared row sets A (n rows) and B (m rows):
 *   output = CGGP_OUT_KERNEL   : k(a_i, b_j) = variance * K_r2(|a_i|^2 + |b_j|^2 - 2 a_i.b_j)   (GPflow expanded form,
 *                                 no clamp for SE, max(r2, 1e-36) before sqrt for Matern)  [+ jitter on i == j]
 *   output = CGGP_OUT_DISTANCE : cggp/distance.py value for `distance` (euclidean uses the difference form of
 *                                 distance.py:9-11 on the prepared rows; sq-euclidean the expanded GPflow form). */
int cggp_kernel_matrix(cggp_ctx* ctx, int dtype, int kind, double variance, int output, int distance,
                       const void* dev_PA, const void* dev_normsA, int64_t n,
                       const void* dev_PB, const void* dev_normsB, int64_t m,
                       int D, int64_t ldp, double jitter, void* dev_out, int64_t ldo);

/* Backward of cggp_kernel_matrix(output = CGGP_OUT_KERNEL) w.r.t. the hyper-parameters - what TensorFlow autodiff
 * computes behind the reference's training loop (cggp/optimize.py:198-254; gradients checked in cggp/cg_test.py:40-46):
 * given dev_G = dL/dK [n, ldg], writes dL/dvariance (1 element) and dL/dlengthscales (D elements; the D contributions
 * are summed by the caller for an isotropic kernel).  Deterministic (fixed-order reduction). */
int cggp_kernel_matrix_backward(cggp_ctx* ctx, int dtype, int kind, double variance,
                                const void* dev_PA, int64_t n, const void* dev_PB, int64_t m, int D, int64_t ldp,
                                const double* host_lengthscales, int ls_count, const void* dev_G, int64_t ldg,
                                void* dev_g_variance, void* dev_g_lengthscales);

/* Nearest-centre assignment (cggp/selection.py:14-32, cggp/optimize.py:50-51): for every prepared data row the
 * argmin over the m centres of `distance` (first minimum wins, as tf.argmin) and that minimal distance. */
int cggp_nearest_center(cggp_ctx* ctx, int dtype, int kind, double variance, int distance,
                        const void* dev_PX, const void* dev_normsX, int64_t n,
                        const void* dev_PZ, const void* dev_normsZ, int64_t m, int D, int64_t ldp,
                        int64_t* dev_idx, void* dev_dist);
/* Cluster statistics (cggp/optimize.py:53-67,88-96): counts[j] = #{i: idx[i] == j}, sums[j] = sum_{idx[i]==j} y[i]. */
int cggp_cluster_stats(cggp_ctx* ctx, int dtype, const int64_t* dev_idx, const void* dev_y, int64_t n, int64_t m,
                       void* dev_counts, void* dev_sums);

/* ---------------------------------------------------------------------------------------------------------
 * Matrix-free product with Kuf Kfu  (the north-star operator's data term)
 *   W[b, :] = V[b, :] @ (Kuf Kfu),  Kfu[i, j] = k(x_i, z_j) never materialised; X is this rank's shard.
 *   variant: 0 = auto, 1 = simple two-sweep kernels (any dtype / D), 3 = fused software-pipelined kernels (TMA-staged
 *   X tiles, K parked in shared memory; the default where supported: float64, D <= 31): every Gram entry is evaluated
 *   once per application; B <= 2 right-hand sides per sweep with FMA contractions, from B = 3 on eight per sweep with
 *   both tile contractions on DMMA (csrc/matvec_pipe8.cu).  (2 was a register-tile kernel of round 1, removed.)
 *   float32 has its own tensor-core entry point, cggp_kuf_kfu_matvec_tf32 below.
 *   The result is NOT all-reduced; call cggp_allreduce_sum (cggp_cg_solve does it per iteration).
 * ------------------------------------------------------------------------------------------------------- */
int cggp_kuf_kfu_matvec(cggp_ctx* ctx, int dtype, int kind, double variance,
                        const void* dev_PX, const void* dev_normsX, int64_t n,
                        const void* dev_PZ, const void* dev_normsZ, int64_t m, int D, int64_t ldp,
                        const void* dev_V, int64_t ldv, int B, void* dev_W, int64_t ldw, int variant);

/* float32 on the 5th-generation tensor cores (tcgen05.mma, FP32 accumulators in TMEM; csrc/matvec_tf32.cu).
 * cggp_tf32_prepare converts prepared float32 points ONCE into the layout the tensor cores read: features padded to
 * KP = cggp_tf32_kp(D) (a multiple of 32), rows padded to cggp_tf32_rows(n) (a multiple of 128), 128-row x 32-feature
 * chunks in the UMMA canonical K-major order, the parts of a chunk interleaved ([row tile][K chunk][part]) so that a
 * tile's operands are one contiguous block = one TMA bulk copy.  `nsplit` selects the arithmetic:
 *   3  -> 3xTF32 (kind::tf32): x = big + small to 2^-22, x.z = b.b + s.b + b.s; float32-accurate distances
 *   1  -> one TF32 pass: 3x fewer tensor-core flops, ~1e-3 relative on the distances
 *   16 -> 3xFP16 (kind::f16): FP16 carries the same 11 significant bits as TF32 at twice the tensor-core rate; a
 *         per-row power-of-two scale supplies the exponent range: x 2^s = H + R (+ 2^-22),
 *         x.z 2^(sx+sz) = H.H + R.H + H.R.  Same accuracy as 3xTF32.
 * Buffers (sizes in floats from cggp_tf32_sizes): dev_stream = the interleaved arrays the products stream (TF32 big
 * [| small]; FP16 H | R), dev_rows = the 1 / row scales of nsplit 16 (1 float otherwise),
 * dev_norms_pad = cggp_tf32_rows(n) floats.  The products take the same nsplit the arrays were prepared with.
 * cggp_kuf_kfu_matvec_tf32: W[B, m] = V[B, m] @ (Kuf Kfu) from those arrays. */
int cggp_tf32_kp(int D);
int64_t cggp_tf32_rows(int64_t n);
int cggp_tf32_sizes(int nsplit, int64_t n, int D, int64_t* stream_floats, int64_t* rows_floats);
/* 1 if the device is sm_100+ and the row tile fits in tensor memory next to the two accumulators
 * (D <= 128: the limit of cggp_prepare_points), else 0 */
int cggp_tf32_supported(cggp_ctx* ctx, int D, int nsplit);
/* host-side query of the kernel's shared-memory ring for feature count D and nb (1 or 2) right-hand sides: K chunks per
 * stage, stages (0 = does not fit), MMA-issuing threads, dynamic shared memory in bytes */
int cggp_tf32_ring_plan(int nsplit, int D, int nb, int* gc, int* stages, int* niss, int64_t* smem_bytes);
int cggp_tf32_prepare(cggp_ctx* ctx, int nsplit, const void* dev_P, const void* dev_norms, int64_t n, int D,
                      int64_t ldp, void* dev_stream, void* dev_rows, void* dev_norms_pad);
int cggp_kuf_kfu_matvec_tf32(cggp_ctx* ctx, int kind, double variance,
                             const void* dev_Xstream, const void* dev_Xrows, const void* dev_xnorms_pad, int64_t n,
                             const void* dev_Zstream, const void* dev_Zrows, const void* dev_znorms_pad, int64_t m,
                             int D, const void* dev_V, int64_t ldv, int B, void* dev_W, int64_t ldw, int nsplit);

/* float32 Kuf @ Y on the tensor cores: W[p, j] = sum_i k(z_j, x_i) Yt[p, i], Y given TRANSPOSED ([P, ldy] row-major) */
int cggp_kuf_times_tf32(cggp_ctx* ctx, int kind, double variance,
                        const void* dev_Xstream, const void* dev_Xrows, const void* dev_xnorms_pad, int64_t n,
                        const void* dev_Zstream, const void* dev_Zrows, const void* dev_znorms_pad, int64_t m,
                        int D, const void* dev_Yt, int64_t ldy, int P, void* dev_W, int64_t ldw, int nsplit);

/* W[p, j] = sum_i k(z_j, x_i) Y[i, p]   (Kuf @ Y over this rank's shard, the right-hand side `Kuf y` of the SGPR
 * system and GPflow's `A @ err`): fused, Kuf never materialised (the second contraction of the pipelined kernel with
 * the row weights given).  Y is [n, ldy] row-major with P columns, W is [P, ldw].  float64, D <= 31; CGGP_ERR_UNSUPPORTED
 * otherwise (the Python layer then forms Kuf in row chunks with cggp_kernel_matrix).  Not all-reduced. */
int cggp_kuf_times(cggp_ctx* ctx, int dtype, int kind, double variance,
                   const void* dev_PX, const void* dev_normsX, int64_t n,
                   const void* dev_PZ, const void* dev_normsZ, int64_t m, int D, int64_t ldp,
                   const void* dev_Y, int64_t ldy, int P, void* dev_W, int64_t ldw);

/* G[m, m] (+)= Kuf Kfu over this rank's n rows (G[a, b] = sum_i k(z_a, x_i) k(x_i, z_b)): the Gram matrix GPflow's
 * SGPR materialises as A A^T (cggp/cli_utils.py:444-446 -> gpflow.models.SGPR), needed where a
 * log-determinant or a trace is (SGPR.elbo) and for Nystrom-style preconditioners.  Kuf is evaluated in L2-sized row
 * chunks and contracted by the library's own FP64 DMMA GEMM as a symmetric rank-k update (lower tiles only, mirrored
 * at the end); float32 uses the FFMA tile GEMM.  accumulate != 0 adds to the G passed in (symmetric).  Not all-reduced. */
int cggp_kuf_gram(cggp_ctx* ctx, int dtype, int kind, double variance,
                  const void* dev_PX, const void* dev_normsX, int64_t n,
                  const void* dev_PZ, const void* dev_normsZ, int64_t m, int D, int64_t ldp,
                  void* dev_G, int64_t ldg, int accumulate);

/* Y[B, n] = V[B, n] @ A[n, n] for SYMMETRIC A (CG's `state.p @ A`, cggp/conjugate_gradient.py:65,74,87). */
int cggp_symm_matmul(cggp_ctx* ctx, int dtype, const void* dev_A, int64_t lda, int64_t n,
                     const void* dev_V, int64_t ldv, int B, void* dev_Y, int64_t ldy);

/* ---------------------------------------------------------------------------------------------------------
 * Conjugate gradient  (replaces cggp/conjugate_gradient.py:24-122, forward pass)
 * ------------------------------------------------------------------------------------------------------- */
enum cggp_operator_type {
  CGGP_OP_DENSE = 0,  /* A given as [n, n] (the reference's only form: Kuu + Lambda, cggp/models.py:301,337) */
  CGGP_OP_SGPR = 1    /* A = Kuu_dense + scale * Kuf Kfu, matrix-free over this rank's X shard, all-reduced */
};

typedef struct cggp_operator {
  uint32_t struct_size; /* sizeof(cggp_operator) of the header the caller was built against: the library rejects any
                           other value instead of reading past a shorter struct (zero-initialise, then set this) */
  int32_t type;
  int32_t dtype;
  int32_t kind;         /* CGGP_OP_SGPR only */
  int64_t n;            /* system size (= M) */
  /* dense part: A (CGGP_OP_DENSE) or Kuu incl. jitter (CGGP_OP_SGPR); symmetric, row-major */
  const void* dev_A;
  int64_t lda;
  /* CGGP_OP_SGPR only */
  int32_t D;
  int32_t variant;      /* as cggp_kuf_kfu_matvec */
  double variance;
  double scale;         /* 1 / noise_variance */
  const void* dev_PX;   /* prepared X shard [n_local, ldp] */
  const void* dev_normsX;
  int64_t n_local;
  const void* dev_PZ;   /* prepared Z [n, ldp] */
  const void* dev_normsZ;
  int64_t ldp;
  /* float32 tensor-core path (optional, CGGP_OP_SGPR with dtype CGGP_F32): arrays from cggp_tf32_prepare; when
   * dev_X32_big is non-NULL every application of the operator goes through cggp_kuf_kfu_matvec_tf32 */
  const void* dev_X32_big;
  const void* dev_X32_small;
  const void* dev_x32_norms;
  const void* dev_Z32_big;
  const void* dev_Z32_small;
  const void* dev_z32_norms;
  int32_t tf32_nsplit;  /* 3, 1 or 16: what the arrays were prepared with */
  int32_t _pad;
} cggp_operator;

enum cggp_precond_type {
  CGGP_PRECOND_EYE = 0,    /* EyePreconditioner, cggp/conjugate_gradient.py:131-134 */
  CGGP_PRECOND_BLOCK = 1,  /* BlockPreconditioner (intent of :137-157): z[blk] = A[blk,blk]^-1 r[blk] */
  CGGP_PRECOND_DENSE = 2   /* z = r @ Pinv for a symmetric [n, n] matrix Pinv ~ A^-1 held on the device (Nystrom- /
                              Cholesky-style preconditioner of the matrix-free operator: the reference's protocol
                              `__call__(vec, mat) -> (z, rz)`, :125-128, with the factorisation done once up front) */
};

typedef struct cggp_precond {
  int32_t type;
  int32_t num_blocks;
  int32_t block_size;
  int32_t _pad;
  const int64_t* dev_block_indices; /* [num_blocks, block_size], a partition of 0..n-1 */
  const void* dev_chol;             /* [num_blocks, block_size, block_size] lower Cholesky factors
                                       (cggp_block_cholesky), or NULL for EYE */
  const void* dev_pinv;             /* CGGP_PRECOND_DENSE: symmetric [n, n], row-major */
  int64_t ldpinv;
} cggp_precond;

/* Gather the diagonal blocks A[blk, blk] of a dense symmetric matrix and factorise them (lower Cholesky). */
int cggp_block_cholesky(cggp_ctx* ctx, int dtype, const void* dev_A, int64_t lda, int64_t n,
                        const int64_t* dev_block_indices, int num_blocks, int block_size, void* dev_chol);

/* z[b, blk] = A[blk, blk]^-1 r[b, blk] for every row b and block: the batched triangular solves of the block-Jacobi
 * preconditioner (one warp per block, lane-parallel forward / backward substitution with the factors of
 * cggp_block_cholesky) - the reference's protocol `__call__(vec, mat) -> (z, rz)`, cggp/conjugate_gradient.py:125-128,
 * outside the solve loop (inside cggp_cg_solve the same routine runs in the fused step kernel). */
int cggp_block_precond_apply(cggp_ctx* ctx, int dtype, int B, int64_t n, const void* dev_r,
                             const cggp_precond* precond, void* dev_z);

/* One fused CG update on [B, n] row-major state given pA = p @ A  (cggp/conjugate_gradient.py:66-84, non-reset branch;
 * with `reset` != 0 the caller passes fresh_r = b - v_new @ A ... see cggp_cg_solve).  Exported for unit tests / ncu.
 *   denom = sum p*pA; gamma = rz/denom (0 where denom <= 1e-16); v += gamma p; r -= gamma pA;
 *   (z, rz') = precond(r); p = z + p*rz'/rz (0 where rz <= 1e-16); rz = rz'; half_rr[b] = 0.5 * sum r^2.  */
int cggp_cg_fused_step(cggp_ctx* ctx, int dtype, int B, int64_t n, const void* dev_pA, void* dev_v, void* dev_r,
                       void* dev_p, void* dev_rz, void* dev_half_rr, const cggp_precond* precond);

/* Full solve, reference semantics (row layout: rhs/x0/solution are [B, n], rows = right-hand sides):
 *   continue while any_b(0.5 |r_b|^2 > error_threshold) and i < max_iterations            (:59-62)
 *   residual refresh r = b - v@A, p = z when i % max_steps_cycle == max_steps_cycle - 1    (:71-84)
 *   returns steps (int32) and half_rz[b] = 0.5 * rz_b                                       (:96-98)
 * The loop runs on the device without a host round trip per iteration: iterations are enqueued in chunks of
 * `check_every`; once the device-side convergence flag is set the remaining enqueued kernels are no-ops.
 * dev_x0 may be NULL (zeros).  dev_history (may be NULL): [history_cap, B] receives 0.5|r_b|^2 at every evaluation of
 * the stopping condition (steps + 1 rows).  host_steps receives the step count; this call synchronises the stream. */
int cggp_cg_solve(cggp_ctx* ctx, const cggp_operator* op, const void* dev_rhs, const void* dev_x0, int B,
                  double error_threshold, int max_iterations, int max_steps_cycle, const cggp_precond* precond,
                  int check_every, void* dev_solution, int32_t* host_steps, void* dev_half_rz,
                  void* dev_history, int64_t history_cap);

/* ---------------------------------------------------------------------------------------------------------
 * Model-level chains  (cggp/models.py objective + prediction as single calls; forward values only - the Python
 * mirror keeps the differentiable path for training)
 * ------------------------------------------------------------------------------------------------------- */

/* CGGP.predict_f, cggp/models.py:333-352, for one batch of nb test points given the dense system
 * A = Kuu + Lambda [m, m] (models.py:337) and a = A^-1 pseudo_u [m] (models.py:339):
 *   Knm = K(Xnew, Z) [nb, m] (the reference's Kmn transposed: rows = right-hand sides)     models.py:334
 *   S   = cg(A, Knm), nb right-hand sides, the reference's stopping rule                   models.py:340
 *   var[b]  = variance - sum_m Knm[b, m] S[b, m]                                           models.py:343-345
 *   mean[b] = sum_m Knm[b, m] a[m]                                                         models.py:351
 * Workspaces dev_Knm_work / dev_S_work are caller-owned [nb, m] buffers (nothing is allocated per call beyond the CG
 * state).  host_steps receives the iteration count of the nb-RHS solve; the call synchronises the stream. */
int cggp_predict_f(cggp_ctx* ctx, int dtype, int kind, double variance,
                   const void* dev_PZ, const void* dev_normsZ, int64_t m,
                   const void* dev_Pnew, const void* dev_normsNew, int64_t nb, int D, int64_t ldp,
                   const void* dev_A, int64_t lda, const void* dev_a,
                   double error_threshold, int max_iterations, int max_steps_cycle,
                   void* dev_Knm_work, void* dev_S_work, void* dev_mean, void* dev_var, int32_t* host_steps);

/* Data term of the ELBO, cggp/models.py:131-133 with GPflow's Gaussian likelihood:
 *   out[0] = sum_i [ -1/2 log(2 pi) - 1/2 log(s2) - 1/2 ((y_i - mean_i)^2 + var_i) / s2 ]
 * (the caller scales by num_data / batch and subtracts the KL, models.py:133-134).  Deterministic two-stage sum. */
int cggp_elbo_terms(cggp_ctx* ctx, int dtype, const void* dev_y, const void* dev_mean, const void* dev_var, int64_t n,
                    double noise_variance, void* dev_out);

/* ---------------------------------------------------------------------------------------------------------
 * Cover-tree inducing-point selection - cggp/covertree.py:25-176 (class CoverTree), called by
 * cggp/optimize.py:19-39 (covertree_update_inducing_parameters).  Same results as the reference on the same rows:
 * its greedy order of operations and NumPy's orders of summation are reproduced (csrc/covertree.cu).
 *
 * build: dev_X [n, D] float64 rows (ldx elements apart).  spatial_resolution > 0 selects the number of levels as the
 * reference does (covertree.py:53-55), otherwise num_levels is used; lloyds / voronoi are the reference's switches
 * (its `distance` argument is ignored there, covertree.py:36-45, and has no counterpart here; `plotting` neither).
 * The call synchronises the ctx stream (the tree structure lives on the host); the handle owns device copies of the
 * leaf memberships and must be released with cggp_covertree_destroy.
 *   level_points  <-> [node.point for node in tree.levels[level]]  (-> `centroids` for the last level); host_parent
 *                     (may be NULL) receives each node's parent index in the previous level
 *   leaf_members  <-> node.data of the last level as row numbers: dev_offsets [m + 1], dev_rows [n]
 *   cluster_stats <-> cluster_mean_and_counts (covertree.py:168-176): np.mean(node.data[1]) (nan for an empty leaf)
 *                     and the count per leaf, both as floating point [m]; dev_y [n] targets, ldy elements apart */
typedef struct cggp_covertree cggp_covertree;
int cggp_covertree_build(cggp_ctx* ctx, int dtype, const void* dev_X, int64_t n, int D, int64_t ldx,
                         double spatial_resolution, int num_levels, int lloyds, int voronoi, cggp_covertree** out);
int cggp_covertree_destroy(cggp_covertree* tree);
int cggp_covertree_num_levels(const cggp_covertree* tree);
int64_t cggp_covertree_level_size(const cggp_covertree* tree, int level);
int cggp_covertree_level_radius(const cggp_covertree* tree, int level, double* host_radius);
int cggp_covertree_level_points(cggp_ctx* ctx, const cggp_covertree* tree, int level, void* dev_out, int64_t ldo,
                                int32_t* host_parent);
int cggp_covertree_leaf_members(cggp_ctx* ctx, const cggp_covertree* tree, int64_t* dev_offsets, int64_t* dev_rows);
int cggp_covertree_cluster_stats(cggp_ctx* ctx, const cggp_covertree* tree, int dtype, const void* dev_y, int64_t ldy,
                                 void* dev_means, void* dev_counts);

/* ---------------------------------------------------------------------------------------------------------
 * Micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json lacks:
 * which: 0 = FP64 DFMA, 1 = FP64 DMMA m8n8k4, 2 = FP64 exp with the 32-entry shuffle table (two-RHS kernels),
 *        3 = FP64 sqrt (rsqrt seed + two Newton steps), 4 = DMMA m16n8k4, 5 = the exp of the pipelined kernels
 *        (1024-entry shared-memory table + degree-3 polynomial), 6 = their sqrt (rsqrt seed + one third-order step).
 * Returns giga-ops/s in *host_gops (FMA counted as 2 flop; exp/sqrt as 1 evaluation). */
int cggp_microbench(cggp_ctx* ctx, int which, int iters, double* host_gops);

#ifdef __cplusplus
}
#endif
#endif /* CGGP_B200_H */
