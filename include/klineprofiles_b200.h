/* klineprofiles_b200.h -- C ABI of the B200-native klineprofiles hot path.
 *
 * Drop-in boundary.  The five XSPEC model symbols at the bottom are exactly the symbols the
 * reference exports (`pub export fn ... callconv(.c)`, reference src/xspec-wrapper.zig:425-530)
 * and that xspec/lmodel.dat:1,9,23,30,43 binds as c_kline ... c_kconv10.  Everything prefixed
 * klp_ is new: the reference evaluates one parameter set per call on one CPU thread; this
 * library evaluates batches of independent parameter sets on one B200 per table, and on all
 * the B200s of a box through a klp_group.
 *
 * There is NO CPU fallback: every evaluation entry point returns KLP_ERR_CUDA (or, for the
 * XSPEC symbols, logs and zero-fills like the reference does on error) when no CUDA device is
 * usable.
 */
#ifndef KLINEPROFILES_B200_H
#define KLINEPROFILES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* models: parameter layouts of reference xspec-wrapper.zig:209-322 / xspec/lmodel.dat */
enum {
    KLP_KLINE = 0,   /* {a, incl_deg, eline, rmin, rmax, alpha}                          6  */
    KLP_KLINE5 = 1,  /* {a, incl_deg, eline, rmin, rmax, rcut, alpha, e1..e5}            12 */
    KLP_KCONV = 2,   /* {a, incl_deg, rmin, rmax, alpha}                  (eline := 1)   5  */
    KLP_KCONV5 = 3,  /* {a, incl_deg, rmin, rmax, rcut, alpha, e1..e5}    (eline := 1)   11 */
    KLP_KCONV10 = 4  /* {a, incl_deg, rmin, rmax, rcut, alpha, e1..e10}   (eline := 1)   16 */
};

/* arithmetic type of the line-profile quadrature */
enum {
    KLP_F32 = 0, /* reference-faithful: f32 exactly as the reference instantiates it
                    (LineProfileTable(2, f32), xspec-wrapper.zig:61,347-359), same operation order,
                    no FMA contraction; rebin / normalise / convolution in f64 as in the reference */
    KLP_F64 = 1  /* every f32 of the reference path replaced by f64 */
};

enum {
    KLP_OK = 0,
    KLP_ERR_ARG = 1,      /* bad argument */
    KLP_ERR_CUDA = 2,     /* CUDA runtime failure / no device */
    KLP_ERR_IO = 3,       /* table file malformed (truncated, unsupported TFORM, HDU/frame-count mismatch ...) */
    KLP_ERR_NOMEM = 4,
    KLP_ERR_NOTFOUND = 5, /* table file cannot be opened (the XSPEC symbols then print the reference's KLINE_PROF_DATA_DIR hint) */
    KLP_ERR_NCCL = 6      /* NCCL not loadable or a collective failed (multi-GPU entry points only) */
};

typedef struct klp_table klp_table; /* host + device copy of a transfer-function table, plus workspaces */

/* Message of the last failing klp_* call on this thread. */
const char* klp_last_error(void);

/* Number of f64 parameters per set for a model (lmodel.dat), or -1. */
int klp_model_nparams(int model);

/* CUDA devices visible to this process (0 and KLP_ERR_CUDA when there is none: the library has no CPU path). */
int klp_device_count(int* n);

/* ---- transfer-function table (replaces reference io.readFitsFile, io.zig:53-87, and
 *      LineProfileTable.init, line-profile.zig:233-273) ------------------------------------ */

/* Build a table from host arrays and upload it to `device` (device = -1: host-only table for the
 * index helpers below; evaluation entry points then return KLP_ERR_CUDA).
 *   spins[n_a], mus[n_mu] ascending (HDU2 / HDU3 of the reference FITS file; mus are cos(incl));
 *   radii/gmin/gmax [n_a*n_mu][n_r] (radii descending within a frame);
 *   upper/lower     [n_a*n_mu][n_r][n_g];
 *   frame index = i_a * n_mu + i_mu (line-profile.zig:60-70).  n_g = 30 in the reference
 *   (xspec-wrapper.zig:75); 2 <= n_g <= 64 here. */
int klp_table_create(const float* spins, int n_a, const float* mus, int n_mu, int n_r, int n_g,
                     const float* radii, const float* gmin, const float* gmax, const float* upper,
                     const float* lower, int device, klp_table** out);

/* Load a table file.  Two formats, told apart by their magic:
 *  - FITS with the reference's schema (io.zig:8-87): HDU 2 column 1 spins, HDU 3 column 1
 *    cos(inclination), HDU 4.. one BINTABLE per (a, mu) frame with columns r, gmin, gmax,
 *    upper_f[n_g], lower_f[n_g] (E or D) -- i.e. the shipped kerr-transfer-functions.fits;
 *    read by a dependency-free reader (no cfitsio);
 *  - the flat ".klpt" container written by lineprofiles_b200.tableio.save_table. */
int klp_table_load(const char* path, int device, klp_table** out);

/* A second handle on the SAME device copy of the frames with its own workspaces, streams and options: evaluations on the
 * two handles may overlap (concurrent walkers / fit threads, one handle per thread or stream) without a second copy of the
 * table.  The frames are freed when the last handle that references them is destroyed. */
int klp_table_share(klp_table* t, klp_table** out);

void klp_table_destroy(klp_table* t);

/* dims[4] = {n_a, n_mu, n_r, n_g} */
int klp_table_dims(const klp_table* t, int* dims);
/* What the table holds (diagnostics; this is how the tests check the FITS reader): the two parameter axes
 * (spins[n_a], mus[n_mu]; either may be NULL) and one frame in the library's layout
 * radii[n_r] gmin[n_r] gmax[n_r] upper[n_r][n_g] lower[n_r][n_g] (copied back from the device for
 * device-resident tables). */
int klp_table_parameters(const klp_table* t, float* spins, float* mus);
int klp_table_read_frame(const klp_table* t, size_t frame, float* out);

/* Host-side index logic, mirrors of line-profile.zig:60-70, :72-93, :209-231, :196-207. */
size_t klp_table_index(const klp_table* t, size_t i_a, size_t i_mu);
void klp_find_parameter_indices(const klp_table* t, float a, float mu, size_t* out2);
void klp_find_tables(const klp_table* t, const size_t* param_indices2, size_t* out4);
int klp_parameter_weight(const klp_table* t, int which, float value, size_t index, float* weight); /* 0 => null */

/* ---- batched evaluation (replaces kerr_lin_emisN / kerr_conv_emisN -> integrate_lineprofile /
 *      kerr_convolve, xspec-wrapper.zig:134-205,326-421, for B independent parameter sets) ---- */

/* Host buffers; the library stages through pinned memory, copies host->device, runs the kernels
 * and copies the spectra back.  Every set is evaluated as a fresh process would (radial-grid
 * limits [r_grid0[0], r_grid0[last]] of setup(), xspec-wrapper.zig:84,149-153).
 *   params   [B][klp_model_nparams(model)] f64
 *   energy   [n_flux+1] bin edges, keV, ascending (shared by all sets)
 *   spectrum convolution models only: input spectrum, [n_flux] shared (spectrum_per_set = 0) or
 *            [B][n_flux]; NULL for kline / kline5
 *   out      [B][n_flux] f64 */
int klp_eval_batch(klp_table* t, int model, int precision, long B, const double* params,
                   const double* energy, int n_flux, const double* spectrum, int spectrum_per_set,
                   double* out);

/* Same, all pointers are device pointers on the table's device; work is enqueued on
 * `cuda_stream` (a cudaStream_t; NULL = default stream) and NOT synchronised.
 * Concurrency: a table handle owns one set of device workspaces (fine-grid flux, deposit scratch,
 * per-call constants), so evaluations on ONE handle must not overlap: klp_eval_batch and the XSPEC
 * symbols serialise themselves (a mutex per handle), device-side calls must be ordered by the
 * caller (same stream, or events between streams).  Concurrent walkers use one handle per stream:
 * klp_table_share gives further handles on the same device copy of the frames. */
int klp_eval_batch_device(klp_table* t, int model, int precision, long B, const double* d_params,
                          const double* d_energy, int n_flux, const double* d_spectrum,
                          int spectrum_per_set, double* d_out, void* cuda_stream);

/* Diagnostics for the parity tests: the fine-grid flux (1999 bins, before rebinning;
 * flux_cache of xspec-wrapper.zig:64,156-160) widened to f64, and for convolution models the line
 * profile on the fixed convolution grid (1998 bins, conv_flux of xspec-wrapper.zig:68,197).
 * Either output pointer may be NULL.  Host buffers. */
int klp_eval_debug(klp_table* t, int model, int precision, long B, const double* params,
                   const double* energy, int n_flux, double* fine_out /*[B][1999]*/,
                   double* profile_out /*[B][1998]*/);

/* Stand-alone frame blend (LineProfileTable.interpolate_parameters, line-profile.zig:116-176):
 * for each set (a, incl_deg) writes the blended frame as T = f32/f64, layout
 * radii[n_r] gmin[n_r] gmax[n_r] upper[n_r][n_g] lower[n_r][n_g].  Device pointers.
 * a_incl: [B][2] f64.  out: [B][n_r*(2*n_g+3)] T. */
int klp_blend_frames_device(klp_table* t, int precision, long B, const double* d_a_incl, void* d_out,
                            void* cuda_stream);

/* Per-kernel device time accumulated over every evaluation of this table since timing was enabled
 * (klp_set_timing(t, 1)) or last read: ms[0] setup, ms[1] quadrature, ms[2] rebin, ms[3] convolution;
 * launches[4] likewise.  Reading resets the accumulation.  Enabling timing adds cudaEvent records on
 * the stream the kernels are launched on. */
int klp_set_timing(klp_table* t, int enable);
int klp_get_timing(klp_table* t, double* ms4, long* launches4);

/* Tuning knobs (testing / benchmarking): "quad_threads" (256, 512 or 1024; 0 = default: 1024 in f32, 512 in f64), "splits" (CTAs
 * that share one set, 0 = auto: more than one only when the batch is smaller than the GPU), "conv_mode" (0 = the
 * reference's operation order in the convolution, bit-exact (default); 1 = cumulative-profile formulation, ~4x
 * faster, equal to rounding), "conv_src_smem" (conv_mode 1 on grids of at most 4096 bins: 1 = the per-source pairs {2/(x_i+x_i+1), spectrum_i}
 * are staged in shared memory next to the cumulative table (default; 20 % faster), 0 = they are read through L1; equal to rounding), "coop" (1 = batches
 * much smaller than the GPU use one cooperative k_quad launch whose CTAs share a set's ordered additions and rebin it (default), 0 = separate
 * kernels; same results), "deposit_ring" (1 = keep the quadrature's deposits in an L2-resident ring per CTA; four consumer warps add them in order, in registers, while the others evaluate:
 * 0.15 GB of scratch instead of 1.8 GB and 6x less DRAM traffic, ~5 % slower; default 0; same results), "ring_discard" (deposit ring only: 1 = a consumer discards the 128-byte lines it has added from the L2 (discard.global.L2) so that dead
 * deposits are never written back to DRAM (default), 0 = leave them to the L2's write-back; same results), "chunk" (sets per kernel launch, default 131072; convolution models at most 65535), "stage_mb" (pinned staging per slot of the
 * host-buffer pipeline, default 256), "stage_tail" (sets of that pipeline's final chunk, whose copy-out nothing overlaps; default 4096, 0 = equal chunks), "order" (1 = evaluate the costliest sets of a launch first (default), 0 = in input order), "fast" (1 = allow the branch-free exact-arithmetic
 * sequences of klp_kernels.cuh when the operand-range checks pass (default), 0 = always use the
 * plain IEEE intrinsics; both give the same results). */
int klp_set_option(klp_table* t, const char* key, long value);
/* The staging chunks klp_eval_batch uses for B sets of n_flux bins with the given "stage_mb" / "stage_tail" (sizes[cap]); returns
 * their number, -1 on bad arguments.  Host logic only: needs no device. */
int klp_stage_chunk_plan(long B, int n_flux, long stage_mb, long stage_tail, long* sizes, int cap);

/* What the most recent quadrature launch of this table was: "k_quad<float, 1024, 1> splits=1 ring=0 fast_frames=600/600"
 * (template instantiation = element type, CTA threads, blended frame in shared memory; CTAs per set; deposit-ring mode; how
 * many table frames pass the operand-range check of the branch-free arithmetic).  bench.py compares it with the kernel name
 * of the committed ncu capture before quoting per-instruction figures from that capture. */
int klp_last_launch_info(klp_table* t, char* buf, size_t len);

/* Diagnostics of the small-batch critical path (option "phase_clock" = 1): microseconds since the start of k_quad at which
 * the CTA holding the first work item passed 1 parameter unpacking, 2 radial grid + frame blend, 3 radial plan, 4 bin ranges,
 * 5 stage A (evaluate + deposit), 6 all CTAs of the set arrived, 7 ordered additions, 8 rebin (cooperative mode only for
 * 5 - 9), 9 the moment the last CTA of the set starts the rebin; -1 where a stamp was not taken.  us10[0] is 0. */
int klp_get_phase_times(klp_table* t, double* us10);

/* Self-test of the branch-free exact-arithmetic sequences (klp_kernels.cuh: div_by, div_gen, sqrt_gen)
 * against the IEEE intrinsics on n_total random + adversarial operands drawn from the ranges they are
 * licensed for.  mismatches4: f32 div_by, f64 div_by, f32 div_gen, f32 sqrt_gen -- all must be 0. */
int klp_arith_selftest(klp_table* t, unsigned long long seed, unsigned long long n_total,
                       unsigned long long* mismatches4);

/* Peak-rate microbenchmark on the table's device: dependent-chain-free FMA throughput in
 * GFLOP/s for precision f32 / f64 (denominator of the compute roofline). */
int klp_fma_peak_gflops(klp_table* t, int precision, double* gflops);

/* ---- several GPUs behind the C ABI (BASELINE.json north_star) -----------------
 * The reference is a single-process, single-thread library (xspec-wrapper.zig:425-530 are plain C calls;
 * XSPEC `parallel` forks).  Independent parameter sets are sharded over the GPUs of one box -- every
 * lookup of the reference is pure (Q18), there is no cross-set exchange -- each GPU holding a replica of
 * the table.  A group is formed either by ONE process that drives n GPUs (klp_group_create_local: what a
 * Zig / XSPEC / Julia host needs) or by one process per GPU (klp_group_create_rank: torchrun / mpirun; the
 * 128-byte id comes from klp_nccl_unique_id on one rank and is distributed by the host).  NCCL is loaded
 * with dlopen at first use; without it everything except the device-resident gather still works.
 * A group borrows its tables: destroy the group first. */
typedef struct klp_group klp_group;
int klp_nccl_unique_id(void* id128);
int klp_nccl_version(int* version);
int klp_group_create_local(klp_table* const* tables, int n, klp_group** out);
int klp_group_create_rank(klp_table* table, const void* id128, int rank, int nranks, klp_group** out);
void klp_group_destroy(klp_group* g);
int klp_group_size(const klp_group* g, int* n_local, int* nranks);

/* Host buffers, local groups: klp_eval_batch over all GPUs of the group.  Contiguous blocks of
 * ceil(B / n) sets per GPU, one host thread and one pinned double-buffered pipeline per GPU, every GPU
 * copies its spectra straight into its slice of `out` -- no collective is needed. */
int klp_group_eval_batch(klp_group* g, int model, int precision, long B, const double* params,
                         const double* energy, int n_flux, const double* spectrum, int spectrum_per_set,
                         double* out);

/* Device-resident buffers: member i (i < n_local) evaluates the B_local sets of d_params[i] -- rank r of
 * the group owns the global sets [r * B_local, (r+1) * B_local) -- and EVERY member ends up with all
 * nranks * B_local spectra in d_out_all[i], laid out [rank][set][bin].  The batch is processed in chunks
 * (options "gather_chunk", default 131072 sets, and "gather_tail", default 2048 sets for the final
 * chunk): while chunk k+1 is integrated, the spectra of chunk k travel over NVLink on
 * a second, high-priority stream -- one grouped NCCL operation per chunk (a single ncclAllGather when the
 * batch is one chunk); only the short final chunk is gathered with nothing left to overlap.  Arrays have n_local entries; d_spectrum may be NULL for
 * kline / kline5; cuda_streams may be NULL (the group's own streams; see klp_group_synchronize).
 * Asynchronous: d_out_all[i] is complete in stream order on cuda_streams[i]. */
int klp_group_eval_allgather_device(klp_group* g, int model, int precision, long B_local,
                                    const double* const* d_params, const double* const* d_energy, int n_flux,
                                    const double* const* d_spectrum, int spectrum_per_set,
                                    double* const* d_out_all, void* const* cuda_streams);
int klp_group_synchronize(klp_group* g);
int klp_group_set_option(klp_group* g, const char* key, long value);
int klp_group_last_gather_ops(const klp_group* g);   /* NCCL group operations issued by the last device-resident call */
/* The chunk sizes the device-resident entry would use for B_local sets with the given options (sizes[cap]); returns their number. */
int klp_group_chunk_plan(long B_local, long gather_chunk, long gather_tail, long* sizes, int cap);

/* ---- the XSPEC model surface (reference xspec-wrapper.zig:425-530) -------------------------
 * void f(const double* energy   n_flux+1 bin edges,
 *        int n_flux,
 *        const double* params   layout per model (above),
 *        int spectrum           unused,
 *        double* flux           additive: output; convolution: input spectrum in, result out,
 *        double* flux_variance  unused, never touched,
 *        const char* init       unused)
 * Process-global state like the reference (xspec-wrapper.zig:57-68): on first use the table is
 * loaded from $KLINE_PROF_DATA_DIR (default ".") -- "kerr-transfer-functions.fits" as in the reference
 * (xspec-wrapper.zig:10-11), else "kerr-transfer-functions.klpt" -- or from the
 * table installed with klp_set_default_table().  Errors: message on stderr prefixed
 * "kerrlineprofile: ", "MODEL OUTPUT SET TO ZERO", flux zero-filled (xspec-wrapper.zig:361-372).
 * The radial-grid limit ratchet of the reference (limits only shrink across calls,
 * xspec-wrapper.zig:149-153) is reproduced; klp_legacy_reset() restores a fresh process. */
void kline(const double* energy, int n_flux, const double* params, int spectrum, double* flux,
           double* flux_variance, const char* init);
void kline5(const double* energy, int n_flux, const double* params, int spectrum, double* flux,
            double* flux_variance, const char* init);
void kconv(const double* energy, int n_flux, const double* params, int spectrum, double* flux,
           double* flux_variance, const char* init);
void kconv5(const double* energy, int n_flux, const double* params, int spectrum, double* flux,
            double* flux_variance, const char* init);
void kconv10(const double* energy, int n_flux, const double* params, int spectrum, double* flux,
             double* flux_variance, const char* init);

/* Install `t` (not owned) as the table used by the XSPEC symbols; NULL uninstalls. */
void klp_set_default_table(klp_table* t);
/* Precision used by the XSPEC symbols (default KLP_F32, the reference's). */
void klp_set_legacy_precision(int precision);
/* Forget the radial-limit ratchet (as if the process had just started). */
void klp_legacy_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* KLINEPROFILES_B200_H */
