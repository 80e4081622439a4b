/* =====================================================================================
 * tplanczos.h -- C ABI of the B200-native two-pass Lanczos engine (libtplanczos.so).
 *
 * This is the drop-in boundary for the reference's hot path (lukefleed/two-pass-lanczos).
 * The reference has no FFI: its boundary is the generic Rust surface
 *     solvers::lanczos / solvers::lanczos_two_pass          src/solvers.rs:46-58, 133-145
 *     algorithms::lanczos::lanczos_standard                 src/algorithms/lanczos.rs:55-61
 *     algorithms::lanczos_two_pass::lanczos_pass_one        src/algorithms/lanczos_two_pass.rs:65-70
 *     algorithms::lanczos_two_pass::lanczos_pass_two[_with_basis]   :128-134, :149-155
 *     utils::data_loader::load_kkt_system                   src/utils/data_loader.rs:211-214
 * over faer's `LinOp<f64>` operator trait.  Each entry point below states which of those it
 * replaces; INTEGRATION.md shows the Rust `extern "C"` block + safe wrappers a maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the boundary; every function returns a
 *     tpl_status (0 = OK) and leaves a human-readable message in tpl_last_error_message()
 *     (thread-local).  Messages equal the reference's `Display` strings (src/error.rs:23-57,
 *     src/utils/data_loader.rs:18-42).
 *   - vectors b, x, y, V may live in host memory OR in device memory of the operator's GPU
 *     (detected with cudaPointerGetAttributes); alphas/betas/steps/b_norm are host outputs.
 *   - all arithmetic is f64.  A handle is not thread-safe; distinct handles are independent (they share
 *     nothing but the kernels' shared-memory opt-in, which the library only ever raises).
 *   - breakdown (beta <= 1000*eps) is NOT an error: it shortens steps_taken (mod.rs:206-211).
 *   - there is no CPU fallback: without a usable CUDA device every compute entry point returns
 *     TPL_ERR_CUDA.
 * ===================================================================================== */
#ifndef TPLANCZOS_H
#define TPLANCZOS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tpl_status {
  TPL_OK = 0,
  /* LanczosErrorKind, declaration order of src/error.rs:20-58 */
  TPL_ERR_BREAKDOWN = 1,          /* never produced (as in the reference) */
  TPL_ERR_DIMENSION_MISMATCH = 2, /* never produced by the reference; here: b/operator size misuse */
  TPL_ERR_INPUT = 3,              /* InputError(String): zero vector b */
  TPL_ERR_PARAMETER_MISMATCH = 4, /* ParameterMismatch{param_name, expected, actual} */
  TPL_ERR_EVD = 5,                /* tpl_ftk_exp did not converge */
  TPL_ERR_SOLVER = 6,             /* SolverError: the f(T_k) callback failed */
  TPL_ERR_PANIC = 7,              /* inputs on which the reference panics (k == 0, null pointers) */
  /* DataLoaderError, declaration order of src/utils/data_loader.rs:16-43 */
  TPL_ERR_IO = 101,
  TPL_ERR_PARSE_INT = 102,
  TPL_ERR_PARSE_FLOAT = 103,
  TPL_ERR_PROBLEM_LINE_MISSING = 104,
  TPL_ERR_UNEXPECTED_EOF = 105,
  TPL_ERR_ARC_COUNT_MISMATCH = 106,
  TPL_ERR_SPARSE_CONSTRUCTION = 107,
  TPL_ERR_INVALID_NODE_INDEX = 108,
  TPL_ERR_MALFORMED_ARC_LINE = 109, /* reference panics: `a` line with < 3 tokens (data_loader.rs:118-119) */
  /* device layer */
  TPL_ERR_CUDA = 200,
  TPL_ERR_COMM = 201
} tpl_status;

const char* tpl_last_error_message(void);
const char* tpl_version(void);

/* ------------------------------------------------------------------------------------
 * KKT loader  --  replaces utils::data_loader::load_kkt_system + KKTSystem
 * (src/utils/data_loader.rs:51-58, 68-259).  Host-side C++; keeps the reference's line
 * semantics verbatim, including the `.qfc` "skip m lines, take <= m lines" rule that yields a
 * SHORT or EMPTY D block on files written by qfcgen (data_loader.rs:172-195).
 * ------------------------------------------------------------------------------------ */
typedef struct tpl_kkt tpl_kkt;

int tpl_load_kkt(const char* dmx_path, const char* qfc_path, tpl_kkt** out);
void tpl_kkt_free(tpl_kkt* kkt);
size_t tpl_kkt_num_nodes(const tpl_kkt* kkt); /* KKTSystem.num_nodes */
size_t tpl_kkt_num_arcs(const tpl_kkt* kkt);  /* KKTSystem.num_arcs  */
size_t tpl_kkt_num_costs(const tpl_kkt* kkt); /* how many quadratic costs the .qfc really gave (<= arcs) */
size_t tpl_kkt_nnz(const tpl_kkt* kkt);
/* KKTSystem.a as faer holds it: CSC, n = nodes + arcs, 8-byte indices, rows ascending per column.
 * Pointers are owned by the handle. */
int tpl_kkt_csc(const tpl_kkt* kkt, size_t* n, size_t* nnz, const uint64_t** colptr,
                const uint64_t** rowidx, const double** val);
/* Incidence view (arc j: +1 at tail, -1 at head, 0-based node ids).  *regular is 1 when every `a` line
 * was a plain arc between two distinct nodes and the arc count matches the `p` line, i.e. when the
 * incidence operator below represents KKTSystem.a exactly. */
int tpl_kkt_incidence(const tpl_kkt* kkt, const uint32_t** tail, const uint32_t** head,
                      const double** d, size_t* d_len, int* regular);

/* Binary instance container (no reference counterpart -- the reference only reads the text pair,
 * data_loader.rs:81-137 line by line).  64-byte header {"TPLKKT1\n", version, nodes, arcs, costs, checksum}, then
 * tail[arcs] and head[arcs] (u32, 0-based, each padded to 8 bytes) and the quadratic costs the .qfc really gave (f64).
 * A loaded container is the same KKTSystem the text pair gives (its CSC is built on first use); saving needs an
 * instance whose incidence view is exact (`regular`). */
int tpl_write_kkt_binary(const char* path, size_t nodes, size_t arcs, const uint32_t* tail, const uint32_t* head,
                         const double* costs, size_t n_costs);
int tpl_kkt_save_binary(const tpl_kkt* kkt, const char* path);
int tpl_load_kkt_binary(const char* path, tpl_kkt** out);

/* ------------------------------------------------------------------------------------
 * Operators  --  the device-resident stand-in for `&impl faer::matrix_free::LinOp<f64>`
 * (src/solvers.rs:56, src/algorithms/mod.rs:167).  Construction uploads the matrix to HBM and builds
 * the kernel-side format; the handle also owns the solver workspace and its CUDA stream.
 * ------------------------------------------------------------------------------------ */
typedef struct tpl_op tpl_op;

/* Generic symmetric sparse operator from a host CSC exactly as SparseColMat<usize,f64> stores it
 * (what `&a.as_ref()` is at src/bin/tradeoff.rs:268).  device < 0: current device. */
int tpl_op_from_csc(size_t n, const uint64_t* colptr, const uint64_t* rowidx, const double* val,
                    int device, tpl_op** out);
/* Dense symmetric operator (the `Mat<f64>` used as `&impl LinOp<f64>` by src/bin/dense_tradeoff.rs:154-162): n x n,
 * column-major with leading dimension lda >= n, copied to the device.  The matrix must be symmetric (as Lanczos
 * requires; the kernels read column i as row i).  SURVEY 8f, N4. */
int tpl_op_from_dense(size_t n, const double* a, size_t lda, int device, tpl_op** out);
/* Dense complex HERMITIAN operator (`T: ComplexField` with T = c64, src/algorithms/mod.rs:167): n x n complex,
 * column-major, every entry stored (re, im), leading dimension lda >= n COMPLEX entries.  The handle's vectors are n complex
 * numbers in the same interleaved storage, i.e. 2 n doubles: tpl_op_nrows() returns 2 n, and every entry point of this header
 * takes and returns such vectors as they are -- with real alpha and beta (T::Real) the recurrence on the interleaved storage IS
 * the complex recurrence (<v, w> = Re(v^H w), ||w||, w - alpha v), only the product A x is complex.  The kernels read column i
 * conjugated as row i.  No test of the reference instantiates a complex operator: parity is checked against a complex128
 * restatement of the same recurrence (tests/test_gpu_dense.py). */
int tpl_op_from_dense_hermitian(size_t n, const double* a, size_t lda, int device, tpl_op** out);
int tpl_op_is_complex(const tpl_op* op); /* 1: vectors are interleaved complex numbers (tpl_op_nrows() / 2 of them) */
/* Diagonal operator diag(d_0 .. d_{n-1}) (the synthetic spectra of src/bin/stability.rs:98-193 and orthogonality.rs): a
 * convenience over tpl_op_from_csc. */
int tpl_op_from_diagonal(size_t n, const double* diag, int device, tpl_op** out);
/* Network-incidence form of A = [[D, E^T], [E, 0]]: reads arc tail/head instead of stored +-1.
 * d_len <= m honours the loader's short-D quirk (rows >= d_len have no diagonal entry). */
int tpl_op_from_kkt(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d,
                    size_t d_len, int device, tpl_op** out);
/* Convenience: operator for a loaded KKTSystem.  format: 0 = auto (incidence when regular, else CSR),
 * 1 = generic CSR, 2 = incidence (error if not regular). */
int tpl_op_from_kkt_system(const tpl_kkt* kkt, int format, int device, tpl_op** out);
void tpl_op_free(tpl_op* op);
size_t tpl_op_nrows(const tpl_op* op); /* LinOp::nrows == ncols */
/* The compute entry points take bare pointers: every vector (b, x, y of apply) must hold tpl_op_nrows(op) doubles.  A
 * binding that knows the length of the buffer it is about to pass calls this first: TPL_ERR_DIMENSION_MISMATCH with the
 * reference's message (src/error.rs:29-35, where faer panics on the mismatch) unless len == nrows. */
int tpl_op_check_len(const tpl_op* op, size_t len);
int tpl_op_format(const tpl_op* op);   /* 1 = CSR, 2 = incidence, 3 = dense */
int tpl_op_device(const tpl_op* op);
/* LinOp::apply: y = A x (used by the reference's property tests, mod.rs:510). */
int tpl_op_apply(tpl_op* op, const double* x, double* y);
/* Use an externally owned CUDA stream (a cudaStream_t passed as void*) instead of the handle's own.
 * STREAM CONTRACT: every copy and kernel of a handle is issued on ITS stream (created non-blocking, i.e. not ordered
 * against the legacy default stream).  A caller that passes DEVICE pointers produced or consumed on another stream must
 * either hand that stream over with this call (what the Python mirror does for device tensors: it adopts the
 * framework's current stream) or order the two streams itself (event / synchronize) before the call and after it.  Host pointers need
 * nothing: the entry points return after the result has arrived. */
int tpl_op_set_stream(tpl_op* op, void* cuda_stream);
/* Device time (ms, CUDA events on the operator's stream) of the last pass-one / pass-two / standard /
 * gemv launched through this handle, and the number of kernels this library launched so far. */
int tpl_op_last_timing(const tpl_op* op, double* pass_one_ms, double* pass_two_ms, double* gemv_ms);
uint64_t tpl_op_kernel_launches(const tpl_op* op);
/* Algorithmic bytes per SpMV of the kernel-side format (B_csr = 12 nnz + 4(n+1),
 * B_inc = 24 m + 4 p) and bytes of HBM the handle currently holds. */
uint64_t tpl_op_matrix_bytes(const tpl_op* op);
uint64_t tpl_op_device_bytes(const tpl_op* op);
/* Execution mode: 0 = automatic (default): persistent cooperative kernels, shared-memory-resident when the
 * per-SM slice of the operator fits (2-D cell partition first, contiguous chunks second), streaming otherwise;
 * 1 = one cooperative launch per Lanczos step (streaming kernels; what a step callback uses); 2 = persistent streaming
 * kernels with tiled node sums even when a resident shape would fit; 3 = streaming kernels with gathered node rows;
 * 4 = chunk-resident kernels even when the cell partition would fit; 5 = blocked streaming kernels (node-block partition,
 * cell-order vectors, bulk-copy input ring) even when a resident shape would fit.  In mode 0 the streaming regime runs
 * the blocked kernels from ~1 M arcs (cells of >= 6.5 k arcs) and the tiled ones below.  All modes run the same per-element arithmetic;
 * they differ in the (fixed) order in which a node row is summed, i.e. by rounding only. */
int tpl_op_set_mode(tpl_op* op, int mode);
/* Name of the kernel family a whole-pass solve through this handle runs: "cells", "chunks", "blocked", "tiled", "gather",
 * "csr", "dense", "sharded" (NCCL phase kernels), "sharded-fused" (tiled kernels spanning all ranks), "sharded-blocked"
 * (blocked kernels spanning all ranks) or "replicated" (sharded handle whose whole operator fits one GPU's on-chip kernels)
 * (static string). */
const char* tpl_op_kernel_shape(const tpl_op* op);
/* Host-only: builds the tile entry lists of the tiled streaming kernels for `ctas` CTAs and tiles of `tile_arcs` arcs on
 * `threads` host threads (0 = automatic) and checks them (every non-loop arc once on its head and once on its tail node
 * per tile, a node's entries of a tile in one thread's slice).  stats = {check code (0 = consistent), tiles per CTA, list
 * entries incl. padding, pieces, padding entries, longest per-thread list, hash of the lists, fold threads}. */
int tpl_tiles_plan(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int ctas, uint32_t tile_arcs, int threads,
                   uint64_t stats[8]);
/* Host-only: builds the blocked streaming layout (2-D node-block partition, cell-order operator, tile lists over local node
 * ids) for a grid of at most `ctas` CTAs with `smem_limit` bytes of shared memory each, on `threads` host threads (0 =
 * automatic), and checks it (gidx a bijection, every packed tail/head word decodes to its arc, cells sorted by tail and
 * padded to whole stages, every non-loop arc once on its tail and once on its head in the lists of its tile).
 * stats = {fits, tail blocks, head blocks, largest tail block, largest head block, padded arcs, tile arcs, tiles of the
 * largest cell, ring slots (pass 1 | pass 2 << 8 | pass 2 with basis << 16), largest cell, smallest cell, list entries incl.
 * padding, pieces, hash of the layout, check code (0 = consistent), shared-memory bytes (pass 2)}. */
int tpl_blocks_plan(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len, int ctas,
                    size_t smem_limit, int threads, uint64_t stats[16]);
/* Diagnostics: the tables of a KKT handle as they sit in device memory, checked on the host.  Handles of 2^20 arcs and more
 * build their node lists and their blocked layout ON THE DEVICE (TPL_HOST_BUILD=1 / TPL_DEVICE_BUILD=1 in the environment
 * force the host / the device builder for any size); this call downloads them, runs the host checker of the blocked layout
 * over them and compares the node lists with a host construction.
 * stats = {built on the device, blocked layout present, check code of the blocked layout (0 = consistent), hash of the
 * layout (the one tpl_blocks_plan reports for the host builder), list words, tile arcs, node-list words that differ,
 * node-list words}. */
int tpl_op_layout_check(tpl_op* op, const uint32_t* tail, const uint32_t* head, const double* d, size_t d_len, uint64_t stats[8]);
/* Diagnostics (host only, no device needed): builds the 2-D cell partition the resident kernels would use on a grid
 * of `ctas` CTAs with `smem_limit` bytes of shared memory each and checks its tables on the host.
 * stats = {fits, tail blocks, head blocks, arc slots per cell, node lines, most entry rows, most node-sum groups,
 * most touched lines, most pushed lines, owned lines per CTA, inbox atoms, largest cell, smallest cell,
 * shared-memory bytes (pass 2), consistency code (0 = consistent), shared-memory wavefronts per half-warp gather x1000
 * (arc rows in the low, node sums in the high 32 bits)}. */
int tpl_cells_plan(size_t m, size_t p, const uint32_t* tail, const uint32_t* head, int ctas, size_t smem_limit,
                   uint64_t stats[16]);
/* Diagnostics: per-CTA, per-step phase timestamps (SM clock) of the resident kernels.  enable(max_steps > 0)
 * allocates ctas x max_steps x marks 64-bit words in HBM, enable(0) frees them; read() copies them out (row-major
 * [cta][step][mark]) and clears the buffer.  Off by default; costs one predicated store per mark when on. */
int tpl_op_trace_enable(tpl_op* op, size_t max_steps);
int tpl_op_trace_read(tpl_op* op, uint64_t* out, size_t capacity, size_t* ctas, size_t* steps, size_t* marks);

/* ------------------------------------------------------------------------------------
 * algorithms::*  building blocks
 * ------------------------------------------------------------------------------------ */

/* lanczos_pass_one (src/algorithms/lanczos_two_pass.rs:65-110).
 * alphas: capacity k, betas: capacity k-1 (may be NULL when k == 1).
 * On return *steps == alphas.len(), betas.len() == *steps - 1, *b_norm == ||b||_2.
 * k == 0 -> TPL_ERR_PANIC (the reference panics in Vec::with_capacity(k-1)). */
int tpl_pass_one(tpl_op* op, const double* b, size_t k, double* alphas, double* betas, size_t* steps,
                 double* b_norm);

/* lanczos_pass_two / lanczos_pass_two_with_basis (lanczos_two_pass.rs:128-166, 206-312).
 * y has y_len entries (must equal steps, else ParameterMismatch{"y_k"}).  x: n outputs.
 * V: NULL, or an n x steps column-major buffer with leading dimension ldv >= n that receives the
 * regenerated basis V'_k (the `_with_basis` variant). */
int tpl_pass_two(tpl_op* op, const double* b, const double* alphas, const double* betas, size_t steps,
                 double b_norm, const double* y, size_t y_len, double* x, double* V, size_t ldv);

/* LanczosCallback (src/algorithms/mod.rs:82-86): called after each step with the step count, the
 * n x steps basis so far (DEVICE pointer, column-major, leading dimension ld), and the T_k view
 * (alphas[steps], betas[steps-1], host).  Return non-zero to continue, 0 to stop. */
typedef int (*tpl_step_callback)(size_t steps, const double* v_dev, size_t ld, const double* alphas,
                                 const double* betas, void* user);

/* lanczos_standard (src/algorithms/lanczos.rs:55-156).  V: n x k column-major, ldv >= n, host or
 * device; columns >= *steps are zero (the reference trims them, lanczos.rs:135-145). */
int tpl_standard(tpl_op* op, const double* b, size_t k, double* V, size_t ldv, double* alphas,
                 double* betas, size_t* steps, double* b_norm, tpl_step_callback cb, void* user);

/* ------------------------------------------------------------------------------------
 * solvers::*  (src/solvers.rs)
 * f_tk_solver closure: F: FnMut(&[R], &[R]) -> Result<Mat<T>, anyhow::Error>  (solvers.rs:58).
 * The callback writes y' into y (capacity *y_len == na on entry) and sets *y_len to the number of
 * rows it produced; a non-zero return becomes SolverError, *y_len != steps becomes
 * ParameterMismatch{"y_k_prime"} (solvers.rs:71-87, 155-165).
 * ------------------------------------------------------------------------------------ */
typedef int (*tpl_ftk_solver)(const double* alphas, size_t na, const double* betas, size_t nb, double* y,
                              size_t* y_len, void* user);

int tpl_lanczos(tpl_op* op, const double* b, size_t k, tpl_ftk_solver f_tk, void* user, double* x);
int tpl_lanczos_two_pass(tpl_op* op, const double* b, size_t k, tpl_ftk_solver f_tk, void* user,
                         double* x);

/* Host f(T_k) e1 solvers with the tpl_ftk_solver signature (the closures the reference's benches
 * and tests define): inverse via tridiagonal partial-pivot LU (src/bin/tradeoff.rs:245-258),
 * exponential via symmetric tridiagonal EVD (src/bin/stability.rs:175-193), square = T*T*e1
 * (tests/correctness.rs:290-299).  `user` is ignored. */
int tpl_ftk_inv(const double* alphas, size_t na, const double* betas, size_t nb, double* y, size_t* y_len,
                void* user);
int tpl_ftk_exp(const double* alphas, size_t na, const double* betas, size_t nb, double* y, size_t* y_len,
                void* user);
int tpl_ftk_square(const double* alphas, size_t na, const double* betas, size_t nb, double* y,
                   size_t* y_len, void* user);

/* k-sweeps (the reference's benches re-solve for every k, src/bin/tradeoff.rs:262-290, src/bin/stability.rs:259-312):
 * x_q = f(A) b with ks[q] Lanczos steps for q < nk from ONE basis generation (tpl_lanczos_sweep: one-pass, the n x max(ks)
 * basis stays in HBM and every x_q comes out of one streaming pass over it, 16 at a time) or ONE pass 1 (tpl_lanczos_two_pass_sweep:
 * O(n) memory, one pass 2 per k).  X: n x nk column-major, leading dimension ldx >= n, host or device.  f_tk is called once per
 * k with the leading ks[q] coefficients.  Every x_q is bit-identical to the corresponding tpl_lanczos / tpl_lanczos_two_pass
 * solve (step j of the recurrence does not depend on k).  ks[q] == 0 -> TPL_ERR_PANIC. */
int tpl_lanczos_sweep(tpl_op* op, const double* b, const size_t* ks, size_t nk, tpl_ftk_solver f_tk, void* user, double* X,
                      size_t ldx);
int tpl_lanczos_two_pass_sweep(tpl_op* op, const double* b, const size_t* ks, size_t nk, tpl_ftk_solver f_tk, void* user,
                               double* X, size_t ldx);

/* SURVEY 8f N1 (no reference counterpart; the reference's benches re-solve for every k, src/bin/tradeoff.rs:262-290):
 * residual norms ||b - A x_j||, j = 1..na, of the iterates x_j = ||b|| V_j T_j^{-1} e_1 from the coefficients of ONE
 * pass 1 (progressive Givens QR of T, O(na), host).  betas[j-1] = beta_j must be given for every j that is wanted
 * (res[j-1] = NaN beyond nb); exact in exact arithmetic, and in floating point up to the loss-of-orthogonality horizon. */
int tpl_ftk_inv_residuals(const double* alphas, size_t na, const double* betas, size_t nb, double b_norm, double* res);
/* A x = b with k chosen from those estimates: pass 1 runs k_max + 1 steps, k = the first j <= k_max whose estimate is
 * <= rtol * ||b|| (else the j with the smallest one), y = T_k^{-1} e_1 ||b||, pass 2 regenerates only k vectors.
 * *k_used and *res_est (estimate at k, absolute) are optional outputs. */
int tpl_lanczos_two_pass_inv_adaptive(tpl_op* op, const double* b, size_t k_max, double rtol, double* x, size_t* k_used,
                                      double* res_est);

/* ------------------------------------------------------------------------------------
 * Multi-GPU (arc-partitioned KKT operator): rank r of `world` owns arcs
 * [arc_begin, arc_end) and a replica of the p node rows.  The communicator is NCCL; the caller
 * supplies the 128-byte ncclUniqueId made by rank 0 (tpl_comm_unique_id) through its own
 * rendezvous (any out-of-band channel; bench.py broadcasts it between its ranks).
 * ------------------------------------------------------------------------------------ */
/* REPLICATED execution: when the WHOLE operator fits the on-chip cell kernels of one GPU (about 600 k arcs), a sharded handle
 * with world > 1 solves the full problem on every rank -- one all-reduce assembles b from the ranks' slices, the passes run
 * without any per-step communication, every rank returns its slice (kernel shape "replicated").  Same rank-local API and
 * results; sharding such a job costs two cross-GPU barriers per Lanczos step and is several times slower than one GPU.
 * tpl_op_set_mode(op, != 0) or the environment variable TPL_NO_REPLICATE=1 (read at construction) keep the arc-partitioned
 * paths; a replicated handle has no exchange block (tpl_op_fabric_export fails, the ranks stay unconnected). */
int tpl_comm_unique_id(uint8_t id_out[128]);
int tpl_op_from_kkt_sharded(size_t m, size_t p, size_t arc_begin, size_t arc_end, const uint32_t* tail,
                            const uint32_t* head, const double* d, size_t d_len, int device, int rank,
                            int world, const uint8_t nccl_id[128], tpl_op** out);
/* Fused multi-GPU exchange (one process per GPU of one NVLink domain): every rank exports the CUDA IPC handle of its
 * exchange block (64 bytes), the caller gathers the handles of all ranks in rank order (any transport) and every rank
 * imports them.  From then on the passes of a sharded operator run as ONE persistent kernel per rank whose node-sum
 * exchange, node-value broadcast and alpha / beta all-reduces are peer-memory stores and local polls inside the kernel
 * (no NCCL call, no per-step launch); tpl_op_set_mode(op, 1) returns to the NCCL phase kernels. */
int tpl_op_fabric_export(tpl_op* op, uint8_t handle[64]);
int tpl_op_fabric_import(tpl_op* op, const uint8_t* handles, int count);
/* rank / world of a handle (0 / 1 for an unsharded one), its local arc count and the node count.  Vectors of a
 * sharded handle are rank-local: [the rank's arc slice | all p node entries (replicated)], nrows() = local_arcs + p. */
int tpl_op_shard_info(const tpl_op* op, int* rank, int* world, size_t* local_arcs, size_t* nodes);

#ifdef __cplusplus
}
#endif
#endif /* TPLANCZOS_H */
