/* forte2_b200 — C ABI of the B200-native determinant-CI sigma builder.
 *
 * This is the boundary a forte2 maintainer binds instead of the nanobind classes
 *   forte2.lib.ci_helpers.CIStrings       (forte2/api/ci_helpers_api.cc:52-85)
 *   forte2.lib.ci_helpers.CISigmaBuilder  (forte2/api/ci_helpers_api.cc:88-187, hot-path subset)
 * Plain pointers and sizes only; no exceptions cross it.  Every function that can fail returns an
 * int status (0 = ok) or a NULL handle / negative count, and leaves a message in f2_last_error().
 * The Python mirror in forte2_b200/ maps the status codes to the exception types nanobind gives the
 * reference: 1 -> ValueError (std::invalid_argument), 2 -> RuntimeError (std::runtime_error),
 * 3 -> OverflowError (std::overflow_error), 4 -> IndexError (std::out_of_range), 5 -> RuntimeError
 * carrying a CUDA error string.
 */
#ifndef FORTE2_B200_H
#define FORTE2_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct f2_strings f2_strings;
typedef struct f2_sigma f2_sigma;
typedef struct f2_adapter f2_adapter;

enum { F2_OK = 0, F2_EINVAL = 1, F2_ERUNTIME = 2, F2_EOVERFLOW = 3, F2_EINDEX = 4, F2_ECUDA = 5 };

/* kinds understood by f2_strings_export (string tables and replacement lists) */
enum {
    F2_STR_N = 0, F2_STR_1H = 1, F2_STR_2H = 2, F2_STR_3H = 3, F2_STR_1H1P = 4,
    F2_VO = 10, F2_VO2 = 11, F2_1H = 12, F2_1H2 = 13, F2_2H = 14, F2_3H = 15
};

const char* f2_last_error(void);
/* status code of the failure f2_last_error() describes (for the functions that return a NULL handle) */
int f2_last_status(void);
int f2_abi_version(void);

/* ---- CIStrings: forte2/ci/ci_strings.h:15, ctor forte2/ci/ci_strings.cc:26-184 ------------------
 * orbital_symmetry is passed flattened: ngas spaces of gas_sizes[g] irrep labels each. */
f2_strings* f2_strings_create(int64_t na, int64_t nb, int symmetry, int ngas, const int* gas_sizes,
                              const int* orbsym_flat, int n_min, const int* gas_min, int n_max,
                              const int* gas_max);
void f2_strings_destroy(f2_strings* h);
/* out[0..11] = na nb nas nbs ndet norb nirrep ngas_spaces nblocks symmetry nclasses_a nclasses_b
 * (properties of forte2/api/ci_helpers_api.cc:71-81) */
int f2_strings_info(const f2_strings* h, int64_t* out);
/* determinant blocks, rows (class_Ia, class_Ib, offset, nIa, nIb); CI-vector layout of
 * forte2/ci/ci_strings.cc:152-160,196-203 */
int f2_strings_blocks(const f2_strings* h, int64_t* out);
/* which: 0 alpha patterns (rows of 6), 1 beta patterns (rows of 6), 2 allowed pairs (rows of 2);
 * returns the number of int64 (gas_alpha_occupations / gas_beta_occupations / gas_occupations) */
int64_t f2_strings_occupations(const f2_strings* h, int which, int64_t* out, int64_t cap);
/* Flatten a string table or replacement list to int64 in the reference's container order
 * (forte2/ci/ci_strings_makers.cc:75-322).  Returns the int64 count (call with out = NULL to size),
 * negative on error.
 *   strings: [nclasses, strpcls[0..nclasses), then the occupation words class by class]
 *   VO   rows: class_I class_J p q sign I J          J = sign a+_p a_q I
 *   VO2  rows: class_K class_I K sign(+-1,+-2) pq J  K in the 1h1p space, pq = pair_geq(p,q)
 *   1H   rows: class_K add_K class_I sign p I        a_p I = sign K
 *   1H2  rows: class_K class_I K sign p I
 *   2H   rows: class_K add_K class_I sign p q I      a_p a_q I = sign K, p > q
 *   3H   rows: class_K add_K class_I sign p q r I
 * kinds 20..28: the device-format tables as the kernels read them (csrc/strings.hpp DeviceLists; packed
 * entries reinterpreted as int64), for host models of the kernels in tests/; kinds 30 + level / 34 + level (level 0..3): CSR
 * pointer / packed entries of the hole lists of the RDM path (csrc/strings.hpp HoleLists) */
int64_t f2_strings_export(const f2_strings* h, int kind, int spin, int64_t* out, int64_t cap);
/* Build the replacement lists now instead of on first use: device >= 0 with CUDA kernels on that GPU
 * (csrc/lists_device.cu; what f2_sigma_create does on its own), device < 0 with the host builder.  Both give the same
 * arrays (forte2/ci/ci_strings_makers.cc:75-322 is the reference counterpart).  No effect once the lists exist. */
int f2_strings_build_lists(f2_strings* h, int device);
/* CIStrings::make_determinants / determinant / determinant_index (ci_strings.cc:186-230):
 * occupation words (alpha, beta) */
int f2_strings_determinants(const f2_strings* h, uint64_t* out /* 2*ndet */);
int f2_strings_determinant(const f2_strings* h, int64_t address, uint64_t* out2);
int64_t f2_strings_determinant_index(const f2_strings* h, uint64_t a, uint64_t b);

/* ---- CISigmaBuilder: forte2/ci/ci_sigma_builder.h:26, ctor ci_sigma_builder.cc:29-47 -------------
 * H[norb*norb], V[norb^4] are HOST pointers, C-contiguous, V in physicist order <pq|rs>
 * (integral hand-off of forte2/jkbuilder/mointegrals.py:40,84).  device = CUDA ordinal.
 * The strings handle must outlive the builder (same rule as the reference, ci_sigma_builder.h:290). */
f2_sigma* f2_sigma_create(f2_strings* strings, double E, const double* H, const double* V,
                          int log_level, int device);
void f2_sigma_destroy(f2_sigma* h);
/* CISigmaBuilder::set_Hamiltonian (ci_sigma_builder.cc:102-185): new integrals, lists are kept */
int f2_sigma_set_hamiltonian(f2_sigma* h, double E, const double* H, const double* V);
/* "hz" | "harrison-zarrabian" | "kh" | "knowles-handy" (ci_sigma_builder.cc:49-57) */
int f2_sigma_set_algorithm(f2_sigma* h, const char* name);
const char* f2_sigma_get_algorithm(const f2_sigma* h);
/* scratch budget in MB (ci_sigma_builder.cc:70-78) */
int f2_sigma_set_memory(f2_sigma* h, int64_t mb);
int f2_sigma_set_log_level(f2_sigma* h, int level);
/* CISigmaBuilder::Hamiltonian (ci_sigma_builder.cc:187-218) on DEVICE buffers of ndet doubles,
 * sigma overwritten; stream = cudaStream_t (NULL = default stream). */
int f2_sigma_apply(f2_sigma* h, const double* C_dev, double* sigma_dev, int64_t n, void* stream);
/* the same for nvec trial vectors, vector v at C_dev + v * ld (ld >= n): what the Davidson-Liu closure of
 * forte2/ci/ci.py:228-233 does column by column, as one call */
int f2_sigma_apply_batch(f2_sigma* h, const double* C_dev, double* sigma_dev, int64_t n, int nvec, int64_t ld,
                         void* stream);
/* same call on HOST buffers (what the reference's numpy-facing method receives): H2D copy, device
 * build, D2H copy, synchronous. */
int f2_sigma_apply_host(f2_sigma* h, const double* C_host, double* sigma_host, int64_t n);
/* avg_build_time (ci_sigma_builder.h:93-102): seconds per build [total, aabb, aaaa, bbbb] */
int f2_sigma_times(f2_sigma* h, double* out4);
/* per-term device time of the LAST build in ms: [H0, same-spin alpha, same-spin beta (with
 * transposes), opposite spin, total] */
int f2_sigma_last_ms(f2_sigma* h, double* out5);
/* determinant energies <I|H|I> (SlaterRules::energy, forte2/determinant/slater_rules.cc:86-121)
 * for every determinant in CI-vector order, on the device: the determinant-basis Davidson
 * preconditioner. */
int f2_sigma_hdiag(f2_sigma* h, double* out_dev, void* stream);
/* number of kernels of this library launched by the last build (bench.py's gpu_launches) */
int64_t f2_sigma_launch_count(const f2_sigma* h);
/* what the opposite-spin kernel executes per build: out[0] = rows M of the integral matrix,
 * out[1] = M padded to the MMA tile, out[2] = max single replacements per alpha string (GEMM K),
 * out[3] = sum over launched (alpha string, beta class) work items of K_item * columns, i.e. the
 * number of (k, column) pairs contracted; executed flops = 2 * out[1] * out[3], useful flops =
 * 2 * out[0] * out[3].  out[4] = 1 when the pair-compressed (8-fold symmetric) rows are in use.
 * out[5] = 1 when the single-GPU build uses the gather form (t3g_kernel: one CTA per input alpha string,
 * rows of C staged in shared memory), 0 for the scatter form (t3_kernel).  out has 6 entries. */
int f2_sigma_t3_shape(const f2_sigma* h, int64_t* out6);

/* ---- alpha-row partition for multi-GPU (no reference counterpart) ----------
 * Rank r of world owns, in every determinant block, the alpha rows [r*nIa/world, (r+1)*nIa/world); its
 * local vector is the concatenation of the owned row slabs in block order.  For the alpha same-spin
 * term (which couples rows) it owns the same kind of slice of every block's beta columns.
 * sizes2 = {local vector length, exchange buffer length}; rows6 = per block
 * (first_row, nrows, local_offset, first_col, ncols, exchange_offset). */
int f2_sigma_part_layout(f2_sigma* h, int rank, int world, int64_t* sizes2, int64_t* rows6);
/* sigma_local (owned rows) from the FULL trial vector (all-gathered by the caller over NCCL):
 * H0 + beta same-spin + opposite-spin are complete on return; the alpha same-spin term of the owned
 * COLUMNS of every row is left in exch_dev, laid out per block as [nIa][ncols], for the caller to
 * send to the row owners (all-to-all) and add.  Device pointers; stream = cudaStream_t. */
int f2_sigma_apply_part(f2_sigma* h, const double* C_full_dev, double* sigma_local_dev, double* exch_dev,
                        int rank, int world, void* stream);

/* Reduced density matrices <L| ... |R> of two device CI vectors (ndet doubles each), one GPU; out_dev is
 * overwritten.  kind 0 a_1rdm, 1 b_1rdm: norb^2 doubles, [p][q] = <L|a+_p a_q|R>
 * (CISigmaBuilder::compute_a_1rdm / compute_b_1rdm, forte2/ci/ci_sigma_builder_1rdm.cc:9-69);
 * kind 3 aa_2rdm, 4 bb_2rdm: npairs^2 doubles, npairs = norb (norb - 1) / 2, [pq][rs] = <L|a+_p a+_q a_s a_r|R>
 * with p > q, r > s and pair index p (p - 1) / 2 + q (compute_aa_2rdm / compute_bb_2rdm,
 * ci_sigma_builder_2rdm.cc:45-133); kind 5 ab_2rdm: norb^4 doubles, [u][v][x][y] =
 * <L|a+_u(alpha) a+_v(beta) a_y(beta) a_x(alpha)|R> (compute_ab_2rdm, ci_sigma_builder_2rdm.cc:135-201).
 * The spin-free combinations (sf_1rdm, sf_2rdm, sf_2cumulant; ci_sigma_builder_1rdm.cc:71-81,
 * ci_sigma_builder_2rdm.cc:203-279) are assembled from these by the Python mirror. */
int f2_sigma_rdm(f2_sigma* h, int kind, const double* L_dev, const double* R_dev, double* out_dev, void* stream);
/* The general form behind it, every spin block up to three-body: the matrix
 *   out[opA * Pb + opB][opA' * Pb + opB'] = <L| (la alpha creators A)(lb beta creators B)(B' annihilators)(A' annihilators) |R>
 * over decreasing orbital tuples (index p; p (p-1)/2 + q; p (p-1) (p-2)/6 + q (q-1)/2 + r), Pb = C(norb, lb), 1 <= la + lb <= 3.
 * (1,0) a_1rdm, (0,1) b_1rdm, (2,0) aa_2rdm, (1,1) ab_2rdm [u][v][x][y], (0,2) bb_2rdm, (3,0) aaa_3rdm, (2,1) aab_3rdm
 * [(uv)][w][(xy)][z], (1,2) abb_3rdm, (0,3) bbb_3rdm (compute_*_3rdm, forte2/ci/ci_sigma_builder_3rdm.cc:86-318; f2_sigma_rdm
 * kinds 8..11).  Gather + FP64 tensor-core GEMM over chunks of the hole index sized by f2_sigma_set_memory, as the reference's
 * K blocks are (ci_sigma_builder_3rdm.cc:235-258); the same "too small for one K-block column" error. */
int f2_sigma_rdm_block(f2_sigma* h, int la, int lb, const double* L_dev, const double* R_dev, double* out_dev,
                       void* stream);
/* One part of nparts of an RDM block (kinds of f2_sigma_rdm): the contributions of a slice of the hole strings (of the input
 * alpha strings for the gather-form opposite-spin block).  The blocks of all parts add up to the block: several GPUs that
 * each hold full copies of the two vectors compute one part each and all-reduce (forte2_b200/dist.py sharded_rdm; the
 * state-averaged RDMs of a sharded solve, forte2/base_classes/ci_base.py:104-149). */
int f2_sigma_rdm_part(f2_sigma* h, int kind, const double* L_dev, const double* R_dev, double* out_dev, int part, int nparts,
                      void* stream);
/* transition 1-RDM <L| a+_p a_q |R> of spin 0 (alpha) / 1 (beta) between the CI spaces of two builders with the same orbitals
 * and electron numbers (different symmetry or GAS restrictions): norb x norb doubles (CISigmaBuilder::compute_a_1trdm /
 * compute_b_1trdm, forte2/ci/ci_sigma_builder_1trdm.cc:11-118; same error messages).  Synchronises the stream. */
int f2_sigma_trdm1(f2_sigma* h_left, f2_sigma* h_right, int spin, const double* L_dev, const double* R_dev,
                   double* out_dev, void* stream);

/* determinant energies of the owned rows, in the local (row slab) layout; synchronises the stream */
int f2_sigma_hdiag_part(f2_sigma* h, double* out_local_dev, int rank, int world, void* stream);
/* the same averaged over the determinants of each orbital configuration (same doubly / singly occupied orbitals):
 * a diagonal that commutes with S^2, the preconditioner of determinant-basis solves with spin control where the
 * reference would use form_Hdiag_csf (ci_sigma_builder.cc:317-370) through its CSF adapter */
int f2_sigma_hdiag_spin_averaged(f2_sigma* h, double* out_local_dev, int rank, int world, void* stream);

/* Rows form of the sharded build: no rank holds the full trial vector.  Available when the CI space is
 * served by a gather-form opposite-spin kernel (f2_sigma_t3_shape out[5]): one determinant block, or
 * symmetry / GAS blocks whose slot layout fits 32-bit offsets.
 * out4 = {available (0/1), slot, nbs, column-slab length}: the reduce-scatter buffer has world slots of
 * `slot` doubles (the longest local vector of any rank), slot q laid out like the local vector of rank
 * q (owned row slabs of every block, in block order). */
int f2_sigma_rows_form(f2_sigma* h, int rank, int world, int64_t* out4);
/* C_local: owned rows of C.  xcols: per block [nIa][owned columns] of C (the layout of the exchange
 * buffer of f2_sigma_apply_part), assembled by the caller with an all-to-all of the row slabs.  On return sigma_local = H0 + beta same-spin term of the owned rows;
 * ycols = alpha same-spin term of [nIa][owned columns] (caller: all-to-all back to the row owners, add);
 * scatter = opposite-spin contributions of the owned INPUT rows to every sigma row in the slot layout
 * above (caller: reduce-scatter over NCCL, add).  phases: 1 = only the opposite-spin term (launched
 * first so that the reduce-scatter can overlap the rest; reads C_local only, so the all-to-all that assembles xcols can
 * still be in flight), 2 = only H0 and the same-spin terms, 3 = both; 2 splits into 4 = H0 and the alpha same-spin term
 * (reads xcols, writes ycols: the all-to-all back can start when it is done) and 8 = the beta same-spin term (reads C_local).
 * A build is phase 1, then 2 (or 4, then 8).  Device pointers; stream = cudaStream_t. */
int f2_sigma_apply_rows(f2_sigma* h, const double* C_local_dev, const double* xcols_dev,
                        double* sigma_local_dev, double* ycols_dev, double* scatter_dev, int rank, int world,
                        int phases, void* stream);
/* Phase 1 of f2_sigma_apply_rows in pieces, for spaces with one determinant block: the opposite-spin term of the local
 * input rows [row_first, row_first + row_count) (rows counted inside the rank's slab), so that a shard that lives in host
 * memory is multiplied while its later rows are still being copied.  flags: 1 = first piece of a build (scatter is
 * zeroed), 2 = last piece; a build is the pieces in any order covering every local row once, then phase 2 (or 4, 8) of
 * f2_sigma_apply_rows.  No reference counterpart (the reference holds its vectors in host memory and has no device). */
int f2_sigma_apply_rows_t3_piece(f2_sigma* h, const double* C_local_dev, double* scatter_dev, int rank, int world,
                                 int64_t row_first, int64_t row_count, int flags, void* stream);

/* ---- NCCL inside the library (f2_dist_init): the rows form of the sharded build with its collectives issued
 * by the library, for hosts that do not bring torch.distributed.  libnccl.so.2 is resolved at run time (dlopen).
 *   rank 0: f2_dist_unique_id(id) and hands the 128 bytes to every rank (any channel: MPI, a file, torch broadcast);
 *   every rank: f2_dist_init(builder, id, rank, world)   ncclCommInitRank on the builder's device, exchange plans, buffers;
 *   f2_dist_apply(builder, C_local, sigma_local, stream)  one sigma build: sigma_local = (H C)_local, both the rank's row slab
 *                                                         (f2_sigma_part_layout; f2_dist_local_size doubles).  Collective.
 *   f2_dist_allreduce                                     in-place sum over the ranks (subspace reductions of a Davidson loop)
 *   f2_dist_finalize                                      destroys the communicators (also done by f2_sigma_destroy).
 * The all-to-all of C runs beside the opposite-spin kernel, the reduce-scatter beside the same-spin kernels, the all-to-all of
 * the alpha same-spin result beside the beta same-spin kernel (two communicators, two library-owned streams).
 * Fails with F2_ERUNTIME where f2_sigma_rows_form reports the rows form unavailable. */
int f2_dist_unique_id(char* out128);
int f2_dist_init(f2_sigma* h, const char* id128, int rank, int world);
int f2_dist_local_size(f2_sigma* h, int64_t* out);
int f2_dist_apply(f2_sigma* h, const double* C_local_dev, double* sigma_local_dev, void* stream);
int f2_dist_allreduce(f2_sigma* h, double* buf_dev, int64_t count, void* stream);
int f2_dist_finalize(f2_sigma* h);

/* <I|H|J> for the n determinants at the given CI-vector addresses, n x n row-major, on the HOST
 * (SlaterRules::slater_rules, forte2/determinant/slater_rules.cc:123-245): the guess Hamiltonian of
 * forte2/ci/ci.py:411-452.  H, V as in f2_sigma_create.  Needs no GPU. */
int f2_slater_matrix(const f2_strings* h, double E, const double* H, const double* V,
                     const int64_t* addresses, int n, double* out);
/* <I|S^2|J> for the determinants at the given CI-vector addresses, n x n row-major on the host (det.spin2,
 * forte2/determinant/determinant_helpers.cc:154-203; used as forte2/sci/sci.py:405-455 uses it) */
int f2_spin2_matrix(const f2_strings* h, const int64_t* addresses, int n, double* out);

/* ---- CISpinAdapter: forte2/determinant/ci_spin_adapter.h:48, forte2/api/ci_helpers_api.cc:190-201 ----
 * CSF <-> determinant transform of the CI vector, executed before and after every sigma build of a
 * spin-adapted CI (forte2/ci/ci.py:228-233).  f2_adapter_create is prepare_couplings
 * (ci_spin_adapter.cc:137-200) over the determinants of `strings` in CI-vector order and needs no GPU;
 * the coupling table goes to the device on the first transform.  twoMs must equal na - nb. */
f2_adapter* f2_adapter_create(const f2_strings* strings, int twoS, int twoMs);
void f2_adapter_destroy(f2_adapter* h);
/* out4 = nconf, ncsf, ndet, ncoupling (properties of ci_helpers_api.cc:196-197) */
int f2_adapter_info(const f2_adapter* h, int64_t* out4);
/* the table CSF(i) = sum_j coeff_ij Determinant(det_ij) in the reference's order
 * (csf_to_det_bounds_ / csf_to_det_coeff_, ci_spin_adapter.h:189-191): bounds[ncsf+1], det[ncoupling],
 * coeff[ncoupling]; returns ncoupling (det/coeff may be NULL to size) */
int64_t f2_adapter_table(const f2_adapter* h, int64_t* bounds, int64_t* det, double* coeff, int64_t cap);
/* first CSF of every configuration, nconf + 1 entries (conf_to_csfs_range, ci_spin_adapter.h:76) */
int f2_adapter_conf_ranges(const f2_adapter* h, int64_t* out);
/* csf_C_to_det_C / det_C_to_csf_C (ci_spin_adapter.cc:69-97) on DEVICE vectors; the output is
 * overwritten.  device = CUDA ordinal of the buffers, stream = cudaStream_t. */
int f2_adapter_csf_to_det(f2_adapter* h, const double* csf_dev, double* det_dev, int device, void* stream);
int f2_adapter_det_to_csf(f2_adapter* h, const double* det_dev, double* csf_dev, int device, void* stream);
/* the same on HOST vectors (the reference's numpy-facing methods): staged through the device */
int f2_adapter_csf_to_det_host(f2_adapter* h, const double* csf_host, int64_t ncsf, double* det_host, int64_t ndet);
int f2_adapter_det_to_csf_host(f2_adapter* h, const double* det_host, int64_t ndet, double* csf_host, int64_t ncsf);
/* Davidson preconditioner in the CSF basis, H_II ~ sum_d c_dI^2 <d|H|d> (CISigmaBuilder::form_Hdiag_csf
 * with spin_adapt_full_preconditioner = false, ci_sigma_builder.cc:331-367), from the determinant
 * energies of f2_sigma_hdiag; device vectors */
int f2_adapter_hdiag_csf(f2_adapter* h, const double* det_energies_dev, double* out_csf_dev, int device,
                         void* stream);

/* ---- Davidson-Liu device primitives (forte2/helpers/davidsonliu.py:184-484) ----------------------
 * The reference solver is numpy on (size, max_subspace) arrays; these are the streaming passes of
 * its iteration on DEVICE vectors stored as contiguous rows (row i at base + i*ld).  At most 16
 * vectors per side of a call.  work_dev: device scratch of f2_dl_workspace_doubles() doubles.
 * Small coefficient matrices are HOST pointers (row-major), results of reductions are written to
 * DEVICE memory (the caller all-reduces them across ranks when the vectors are sharded). */
int64_t f2_dl_workspace_doubles(void);
/* out_dev[i*q + j] = <A_i | X_j>   (subspace matrix B^T S, davidsonliu.py:257; Gram-Schmidt
 * overlaps, davidsonliu.py:445-468) */
int f2_dl_dots(const double* A, int64_t lda, int p, const double* X, int64_t ldx, int q, int64_t n,
               double* out_dev, double* work_dev, void* stream);
/* Y_j = beta*Y_j + sum_i coef[i*k + j] A_i, j < k; rows of Y may alias rows of A (in-place collapse,
 * davidsonliu.py:389-413; Ritz vectors :376-378; projection against the basis :317).  sumsq_dev
 * (may be NULL) receives |Y_j|^2. */
int f2_dl_combine(const double* A, int64_t lda, int p, const double* coef_host, double* Y, int64_t ldy,
                  int k, double beta, int64_t n, double* sumsq_dev, double* work_dev, void* stream);
/* R_j = sum_i alpha[i*k+j] (S_i - lam_j B_i); rnorm2_dev[j] = |R_j|^2; then the preconditioner
 * R_j <- R_j / (lam_j - hdiag) where |lam_j - hdiag| > denom_threshold, else 0
 * (davidsonliu.py:266-283).  hdiag = NULL skips the preconditioner. */
int f2_dl_residual(const double* B, const double* S, int64_t ld, int m, const double* alpha_host,
                   const double* lam_host, const double* hdiag, double* R, int64_t ldr, int k,
                   double denom_threshold, int64_t n, double* rnorm2_dev, double* work_dev, void* stream);
int f2_dl_scale(double* v, double s, int64_t n, void* stream);

/* Fused passes of one iteration for at most 4 roots (k) and 16 basis vectors (m): the candidate block is
 * orthonormalised as a block (the reference does it vector by vector, davidsonliu.py:415-484).  Each pass reads
 * the basis once and returns the overlaps the next pass needs.
 *   f2_dl_residual_fused  R_j = preconditioned residual as f2_dl_residual;
 *                         out_dev = [ |r_j|^2 (k) ; <B_i|R_j> (m x k) ; <R_j|R_j'> (k x k) ]
 *   f2_dl_project         Z_j = sum_l T[l, j] (Y_l - sum_i c[i, l] B_i), l < kin, j < kout (Z may alias Y; m = 0: no basis)
 *                         out_dev = [ <B_i|Z_j> (m x kout) ; <Z_j|Z_j'> (kout x kout) ] */
int f2_dl_residual_fused(const double* B, const double* S, int64_t ld, int m, const double* alpha_host,
                         const double* lam_host, const double* hdiag, double* R, int64_t ldr, int k,
                         double denom_threshold, int64_t n, double* out_dev, double* work_dev, void* stream);
int f2_dl_project(const double* B, int64_t ld, int m, const double* Yin, int64_t ldin, int kin, const double* c_host,
                  const double* T_host, double* Zout, int64_t ldout, int kout, int64_t n, double* out_dev,
                  double* work_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
