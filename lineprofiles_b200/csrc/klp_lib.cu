// klp_lib.cu -- host side of the C ABI declared in include/klineprofiles_b200.h.
//
// Mirrors the reference's src/xspec-wrapper.zig for the hot path: model dispatch
// (kerr_lin_emisN / kerr_conv_emisN, :326-421), process-global state of the XSPEC symbols
// (:57-68), table lookup by $KLINE_PROF_DATA_DIR (:30-48), log-and-zero error convention
// (:361-372) -- and adds the batched engine on top (chunking, pinned staging, stream overlap).
// No CPU compute path exists in this file: every evaluation launches the kernels of
// klp_kernels.cuh or fails.
#include "../../include/klineprofiles_b200.h"
#include "klp_kernels.cuh"

#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

using namespace klp;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define KLP_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(KLP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    } while (0)

int model_nparams(int model) {
    switch (model) {
        case KLP_KLINE: return 6;
        case KLP_KLINE5: return 12;
        case KLP_KCONV: return 5;
        case KLP_KCONV5: return 11;
        case KLP_KCONV10: return 16;
    }
    return -1;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (dev < 0) { ok = false; return; }  // host-only table
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KLP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail(KLP_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cap = bytes;
        return KLP_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KLP_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) return fail(KLP_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
        cap = bytes;
        return KLP_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Profiler ranges where the reference has its Tracy zone (InterpolatingTransferFunction.integrate,
// transfer-functions.zig:141-142): one range per host call and one per launch group.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// process-global state of the XSPEC symbols (xspec-wrapper.zig:57-68); defined here so that klp_table_destroy can
// uninstall a table that is still referenced
std::recursive_mutex g_legacy_mu;
klp_table* g_default_table = nullptr;  // installed by the caller (not owned)
klp_table* g_owned_table = nullptr;    // lazily loaded singleton (xspec-wrapper.zig:61)

constexpr long long kCtlDoneSets = 2048;   // per-set done flags that fit behind the work counters (see eval_device_t)

struct EvalExtra {
    bool use_limits = false;
    double lim_lo = 0, lim_hi = 0;
    bool want_lims = false;  // write new end points of set 0 into table->d_lims
    bool setup_valid = false; // the per-call constants on the device are those of this very call (host pipeline): skip k_setup
    void* lims_dst = nullptr; // device address for the new end points (default: the table's lims buffer)
};

}  // namespace

// The device copy of the frames, shared by a table and the workspaces cloned from it (klp_table_share).
struct FrameStore {
    int device = 0;
    float* d_frames = nullptr;
    float* d_spins = nullptr;
    float* d_mus = nullptr;
    unsigned char* d_frame_ok = nullptr;
    ~FrameStore() {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        if (d_frames) cudaFree(d_frames);
        if (d_spins) cudaFree(d_spins);
        if (d_mus) cudaFree(d_mus);
        if (d_frame_ok) cudaFree(d_frame_ok);
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct klp_table {
    int device = 0;
    std::shared_ptr<FrameStore> store;   // owner of d_frames / d_spins / d_mus / d_frame_ok once the table is complete
    int n_a = 0, n_mu = 0, n_r = 0, n_g = 0;
    size_t frame_stride = 0;
    std::vector<float> spins, mus;
    std::vector<float> h_frames;   // host-only tables (device == -1) keep their packed frames for klp_table_read_frame
    float* d_frames = nullptr;
    float* d_spins = nullptr;
    float* d_mus = nullptr;
    unsigned char* d_frame_ok = nullptr;   // per frame: values within the binades that license the FAST arithmetic (klp_kernels.cuh)
    size_t frames_fast = 0;                // how many frames pass
    int sm_count = 0;
    int max_smem_optin = 0;
    int max_smem_sm = 0;
    // workspaces
    Buf cc32, cc64, fine, prof, conv_grid, conv_avg, scratch, counter, lims, vals, done, order, phase;
    bool phase_clock = false;   // option "phase_clock": k_quad stamps %globaltimer at its phase boundaries (klp_get_phase_times)
    // host-API staging
    Buf d_energy, d_spec, d_params[2], d_out[2], d_spec_slot[2];
    PinBuf h_params[2], h_out[2], h_spec_slot[2], h_lims;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    // options
    int quad_threads = 0;  // 0 = auto
    bool use_fast = true;  // option "fast"
    int splits = 0;        // 0 = auto
    int conv_mode = 0;     // 0 = the reference's operation order (bit-exact), 1 = cumulative-profile formulation
    int stage_tail = 4096; // sets of the final staging chunk of the host-buffer pipeline (its copy-out overlaps nothing)
    int conv_src_smem = 1; // conv_mode 1: stage the per-source {avg, spectrum} pairs in shared memory (512-thread CTAs; grids up to 4096 bins)
    bool use_order = true;      // evaluate the costliest sets of a launch first
    // the per-call constants currently on the device (host pipeline only): energy grid, precision, conv
    std::vector<double> setup_energy;
    int setup_precision = -1, setup_conv = -1;
    int stage_mb = 256;         // pinned staging per slot of the host pipeline (MB of output)
    bool use_coop = true;       // option "coop": cooperative small-batch mode of k_quad (fused rebin) when the batch is tiny
    bool coop_supported = false;
    bool rebin_attr_set[2] = {false, false};
    std::map<long long, int> quad_occ;   // (k_quad instantiation, shared memory) -> resident CTAs per SM
    bool deposit_ring = false;  // L2 ring of deposits + concurrent ordered additions (no DRAM traffic, -4.5 % throughput)
    bool ring_discard = true;   // deposit ring: consumed lines are discarded from the L2 (discard.global.L2) instead of written back
    bool frame_global = false;  // keep the blended frame in the per-CTA global scratch even when it fits shared memory
    long chunk = 131072;        // sets per kernel launch (convolution models: at most 65535, the gridDim.y limit of their kernels)
    // timing
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<std::pair<int, int>> ev_used;  // (kind, pool index)
    size_t ev_next = 0;
    double ms[4] = {0, 0, 0, 0};
    long launches[4] = {0, 0, 0, 0};
    std::recursive_mutex mu_;   // guards the host-side state of the table (workspaces, options, setup cache, timing events)
    // what the last quadrature launch was (klp_last_launch_info): instantiation and mode
    char quad_variant[96] = "";
};

namespace {

// ---- timing helpers ------------------------------------------------------------------------
struct TimedLaunch {
    klp_table* t;
    cudaStream_t st;
    int kind;
    int idx = -1;
    TimedLaunch(klp_table* t_, cudaStream_t st_, int kind_) : t(t_), st(st_), kind(kind_) {
        if (!t->timing) return;
        if (t->ev_next == t->ev_pool.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            t->ev_pool.push_back({a, b});
        }
        idx = (int)t->ev_next++;
        cudaEventRecord(t->ev_pool[idx].first, st);
    }
    ~TimedLaunch() {
        if (idx < 0) return;
        cudaEventRecord(t->ev_pool[idx].second, st);
        t->ev_used.push_back({kind, idx});
    }
};

template <class T> CallConsts<T>* cc_ptr(klp_table* t);
template <> CallConsts<float>* cc_ptr<float>(klp_table* t) { return (CallConsts<float>*)t->cc32.p; }
template <> CallConsts<double>* cc_ptr<double>(klp_table* t) { return (CallConsts<double>*)t->cc64.p; }

template <class T, int NT, bool FRAME_SMEM> int launch_quad_impl(klp_table* t, QuadArgs<T>& q, size_t smem, size_t fb, cudaStream_t st) {
    // attribute + occupancy of this instantiation: asked once per (instantiation, shared-memory size), not per launch
    const long long key = ((long long)(sizeof(T) == 4 ? 0 : 1) << 60) | ((long long)NT << 40) | ((long long)(FRAME_SMEM ? 1 : 0) << 39) | (long long)smem;
    int occ = 0;
    auto it = t->quad_occ.find(key);
    if (it != t->quad_occ.end()) occ = it->second;
    else {
        KLP_CUDA(cudaFuncSetAttribute(k_quad<T, NT, FRAME_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KLP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_quad<T, NT, FRAME_SMEM>, NT, smem));
        t->quad_occ[key] = occ;
    }
    if (occ < 1) return fail(KLP_ERR_CUDA, "k_quad does not fit on an SM");
    const long long resident = (long long)occ * t->sm_count;
    int splits = t->splits;
    // Cooperative small-batch mode: at least 16 CTAs per set (measured: faster up to ~9 sets on a B200; a single XSPEC call is the extreme).
    // Every CTA of the launch is resident, so the CTAs of a set may wait for each other: items of one radius, the ordered
    // additions shared by bin group, the rebin fused into the last CTA (one kernel instead of two, ~3x shorter critical path).
    bool coop = FRAME_SMEM && t->use_coop && t->coop_supported && splits <= 0 && q.B * 16 <= resident && !t->deposit_ring;
    if (coop) {
        splits = (int)std::min<long long>(resident / q.B, kMaxSplits);
        coop = splits >= 2;
    }
    if (!coop && splits <= 0) {
        // CTAs sharing a set.  Sets differ in cost by a factor of ~2, so a batch of the order of the resident CTAs is
        // balanced by cutting every set into several (set, split) items (each split repeats the plan, ~3 % of a set);
        // measured optimum (tools/splits_probe.py): ~8 items per resident CTA, at most 16 splits (32 for a handful of sets).
        const long long want = (8 * resident + q.B - 1) / q.B;
        const long long cap = q.B * 16 < resident ? 32 : 16;
        splits = (int)std::max<long long>(1, std::min<long long>(want, cap));
        if (splits == 2) splits = 1;   // measured: two splits never pay for the repeated plan once the sets are cost-ordered
    }
    splits = std::max(1, std::min(splits, kMaxSplits));
    q.splits = splits;
    q.coop = coop ? 1 : 0;
    const long long items = q.B * splits;
    const int grid = (int)std::min<long long>(items, resident);
    if (!FRAME_SMEM) {
        int rc = t->scratch.ensure((size_t)grid * fb);
        if (rc) return rc;
        q.scratch = (T*)t->scratch.p;
    }
    // deposit scratch: a ring of kRingRadii radii per CTA when every set is owned by one CTA (it stays in L2), the whole
    // (radius, bin) plane per set when sets are shared by several CTAs (small batches)
    q.ring_mode = t->deposit_ring ? (t->ring_discard ? 2 : 1) : 0;
    q.vals_stride = (splits > 1 || !q.ring_mode) ? (size_t)kRadial * kFineG : (size_t)kRingRadii * kRingStride;   // dense (radius, bin) addressing
    const size_t regions = splits > 1 ? (size_t)std::min<long long>(q.B, kSplitRegions) : (size_t)grid;
    q.regions = (long long)regions;
    int rc = t->vals.ensure(regions * q.vals_stride * sizeof(T));
    if (rc) return rc;
    q.vals = (T*)t->vals.p;
    q.done = nullptr;
    if (splits > 1) {
        if (q.done_prezeroed != nullptr && q.B <= kCtlDoneSets) {
            q.done = q.done_prezeroed;   // single-launch call: zeroed together with the work counter (one memset instead of two)
        } else {
            if ((rc = t->done.ensure(sizeof(unsigned) * 2 * (size_t)q.B))) return rc;
            KLP_CUDA(cudaMemsetAsync(t->done.p, 0, sizeof(unsigned) * 2 * (size_t)q.B, st));
            q.done = (unsigned*)t->done.p;
        }
    }
    if (coop) {
        q.rb.n_stage = (int)std::min<size_t>(fb / sizeof(double), 1u << 20);   // doubles of the dead frame area the fused rebin may use
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_quad<T, NT, FRAME_SMEM>, (const QuadArgs<T>)q);
        if (e != cudaSuccess) {
            // no co-residency guarantee to be had here (MPS limits, a debugger, ...): never again on this handle, separate kernels instead
            cudaGetLastError();
            t->coop_supported = false;
            return launch_quad_impl<T, NT, FRAME_SMEM>(t, q, smem, fb, st);
        }
    } else {
        k_quad<T, NT, FRAME_SMEM><<<grid, NT, smem, st>>>(q);
    }
    KLP_CUDA(cudaGetLastError());
    std::snprintf(t->quad_variant, sizeof(t->quad_variant), "k_quad<%s, %d, %d> splits=%d ring=%d%s coop=%d", sizeof(T) == 4 ? "float" : "double", NT,
                  FRAME_SMEM ? 1 : 0, splits, q.ring_mode ? 1 : 0, q.ring_mode == 2 ? " discard=1" : "", q.coop);
    return KLP_OK;
}

template <class T, int NT> int launch_quad_nt(klp_table* t, QuadArgs<T>& q, cudaStream_t st) {
    const int NW = NT / 32;
    const size_t fixed = QuadSmem<T>::fixed_bytes(t->n_g, NW);
    const size_t fb = QuadSmem<T>::frame_bytes(t->n_r, t->n_g);
    if (fixed > (size_t)t->max_smem_optin) return fail(KLP_ERR_ARG, "table rows too wide for shared memory");
    // The blended frame lives in shared memory when that does not cost a resident CTA (f32: two 512-thread CTAs per SM are
    // budgeted, see quad_min_blocks); otherwise in the per-CTA global scratch -- it is read once per radius (measured -0.7 %).
    const int want = quad_min_blocks<T, NT>();
    const size_t per_sm = (size_t)t->max_smem_sm;
    const bool fits_one = fixed + fb <= (size_t)t->max_smem_optin;
    const bool fits_want = (size_t)want * (fixed + fb + 1024) <= per_sm;
    const bool global_helps = (size_t)want * (fixed + 1024) <= per_sm;
    // a handful of sets cannot fill the SMs anyway: keep the frame in shared memory, the plan (row bisection, gmin/gmax
    // lookups) is then not bound by L2 latency
    const bool tiny = q.B * 4 <= (long long)t->sm_count;
    if (fits_one && !t->frame_global && (fits_want || !global_helps || tiny)) return launch_quad_impl<T, NT, true>(t, q, fixed + fb, fb, st);
    return launch_quad_impl<T, NT, false>(t, q, fixed, fb, st);
}

template <class T> int launch_quad(klp_table* t, QuadArgs<T>& q, cudaStream_t st) {
    int nt = t->quad_threads;
    // measured (tools/nt_probe.py): f32 -- one 1024-thread CTA per SM beats two of 512 at every batch size (+1.5 % at 5e4
    // sets, +6 % at 2000: one plan per SM instead of two, 32 warps either way); f64 -- 1024 threads spill, 512 it is
    if (nt <= 0) nt = sizeof(T) == 4 ? 1024 : 512;
    switch (nt) {
        case 256: return launch_quad_nt<T, 256>(t, q, st);
        case 512: return launch_quad_nt<T, 512>(t, q, st);
        case 768:   // f32 only: 24 warps with up to 85 registers each (tools/nt_probe.py)
            if constexpr (sizeof(T) == 4) return launch_quad_nt<T, 768>(t, q, st);
            else return fail(KLP_ERR_ARG, "quad_threads 768 is an f32 configuration");
        case 1024: return launch_quad_nt<T, 1024>(t, q, st);
    }
    return fail(KLP_ERR_ARG, "quad_threads must be 256, 512, 768 or 1024");
}

TableDev table_dev(const klp_table* t) {
    TableDev d;
    d.frames = t->d_frames;
    d.spins = t->d_spins;
    d.mus = t->d_mus;
    d.frame_stride = t->frame_stride;
    d.n_a = t->n_a;
    d.n_mu = t->n_mu;
    d.n_r = t->n_r;
    d.n_g = t->n_g;
    d.frame_ok = t->d_frame_ok;
    return d;
}

// One batch on one stream, device pointers.  fine_keep / prof_keep: debug taps (device pointers to
// [B][kFineG] T and [B][kConvN-1] f64) -- when given, B must fit one chunk.
template <class T>
int eval_device_t(klp_table* t, int model, long long B, const double* d_params, const double* d_energy, int n_flux,
                  const double* d_spec, int per_set, double* d_out, cudaStream_t st, const EvalExtra& ex,
                  bool stop_after_profile) {
    const int P = model_nparams(model);
    const bool conv = model >= 2;
    Buf& ccb = sizeof(T) == 4 ? t->cc32 : t->cc64;
    int rc;
    if ((rc = ccb.ensure(sizeof(CallConsts<T>)))) return rc;
    // work counters of up to 4096 launches, then (single-launch calls of at most kCtlDoneSets sets) the per-set done flags
    if ((rc = t->counter.ensure(sizeof(unsigned long long) * 4096 + sizeof(unsigned) * 2 * kCtlDoneSets))) return rc;
    if ((rc = t->lims.ensure(sizeof(double) * 2))) return rc;
    if (conv) {
        if ((rc = t->conv_grid.ensure(sizeof(double) * 3 * kConvN))) return rc;  // grid | bin widths | reciprocals
        if ((rc = t->conv_avg.ensure(sizeof(double) * (size_t)n_flux))) return rc;
    }
    const long long chunk = std::max<long long>(1, std::min<long long>(conv ? std::min<long>(t->chunk, 65535) : t->chunk, B));
    if ((rc = t->fine.ensure(sizeof(T) * (size_t)chunk * kFineG))) return rc;
    if (conv && (rc = t->prof.ensure(sizeof(double) * (size_t)chunk * (kConvN - 1)))) return rc;

    if (!ex.setup_valid) {
        TimedLaunch tl(t, st, 0);
        SetupArgs sa;
        sa.energy = d_energy;
        sa.n_flux = n_flux;
        sa.n_g = t->n_g;
        sa.is_conv = conv ? 1 : 0;
        sa.conv_grid = (double*)t->conv_grid.p;
        sa.conv_avg = (double*)t->conv_avg.p;
        sa.conv_bw = (double*)t->conv_grid.p + kConvN;
        sa.conv_ybw = (double*)t->conv_grid.p + 2 * kConvN;
        k_setup<T><<<1, 128, 0, st>>>(sa, cc_ptr<T>(t));
        KLP_CUDA(cudaGetLastError());
    }
    const long long n_chunks = (B + chunk - 1) / chunk;
    if (n_chunks > 4096) return fail(KLP_ERR_ARG, "batch too large for the chunk size");
    unsigned* done_pre = nullptr;
    if (n_chunks == 1 && B <= kCtlDoneSets) {
        // counter[0] and the done flags (placed right behind it for this purpose) in one memset
        done_pre = reinterpret_cast<unsigned*>((unsigned long long*)t->counter.p + 1);
        KLP_CUDA(cudaMemsetAsync(t->counter.p, 0, sizeof(unsigned long long) + sizeof(unsigned) * 2 * (size_t)B, st));
    } else {
        KLP_CUDA(cudaMemsetAsync(t->counter.p, 0, sizeof(unsigned long long) * (size_t)n_chunks, st));
    }
    for (long long c = 0; c < n_chunks; ++c) {
        const long long off = c * chunk;
        const long long nb = std::min(chunk, B - off);
        QuadArgs<T> q;
        q.tab = table_dev(t);
        q.cc = cc_ptr<T>(t);
        q.params = d_params + (size_t)off * P;
        q.model = model;
        q.P = P;
        q.B = nb;
        q.splits = 1;
        q.fine = (T*)t->fine.p;
        q.scratch = nullptr;
        q.allow_fast = t->use_fast ? 1 : 0;
        q.one = 1.0f;
        q.counter = (unsigned long long*)t->counter.p + c;
        q.use_limits = ex.use_limits ? 1 : 0;
        q.lim_lo = (T)ex.lim_lo;
        q.lim_hi = (T)ex.lim_hi;
        q.lims_out = (ex.want_lims && c == 0) ? (ex.lims_dst ? (T*)ex.lims_dst : (T*)t->lims.p) : nullptr;
        q.done_prezeroed = done_pre;
        q.order = nullptr;
        q.phase_ns = nullptr;
        if (t->phase_clock) {
            if ((rc = t->phase.ensure(sizeof(unsigned long long) * 16))) return rc;
            KLP_CUDA(cudaMemsetAsync(t->phase.p, 0, sizeof(unsigned long long) * 16, st));
            q.phase_ns = (unsigned long long*)t->phase.p;
        }
        if (nb >= 512 && t->use_order) {   // costliest sets first (k_order)
            if ((rc = t->order.ensure(sizeof(int) * (size_t)nb))) return rc;
            TimedLaunch tl(t, st, 0);   // accounted with the set-up kernels
            k_order<<<1, 1024, 0, st>>>(q.params, P, (int)nb, (int*)t->order.p);
            KLP_CUDA(cudaGetLastError());
            q.order = (const int*)t->order.p;
        }
        RebinArgs<T> r;
        r.cc = cc_ptr<T>(t);
        r.params = q.params;
        r.model = model;
        r.P = P;
        r.B = nb;
        r.fine = (const T*)t->fine.p;
        if (conv) {
            r.energy = (const double*)t->conv_grid.p;
            r.n_out = kConvN - 1;
            r.out = (double*)t->prof.p;
            r.out_stride = kConvN - 1;
        } else {
            r.energy = d_energy;
            r.n_out = n_flux;
            r.out = d_out + (size_t)off * n_flux;
            r.out_stride = (size_t)n_flux;
        }
        r.caller_edges = conv ? 0 : 1;
        // per-bin values staged in dynamic shared memory, sized to the grid (1000 bins: 8 + 16 KB per CTA, 9 CTAs per SM;
        // the fixed 32 KB of earlier builds allowed 4, and the sequential in-order total of a set hides behind other sets only)
        r.n_stage = r.n_out <= 4096 ? ((r.n_out + 1) & ~1) : 0;
        q.rb = r;       // used by the cooperative small-batch mode, which rebins inside k_quad
        q.coop = 0;
        {
            NvtxRange nvtx_quad("klp integrate (k_quad)");
            TimedLaunch tl(t, st, 1);
            if ((rc = launch_quad<T>(t, q, st))) return rc;
        }
        if (!q.coop) {
            TimedLaunch tl(t, st, 2);
            if (!t->rebin_attr_set[sizeof(T) == 4 ? 0 : 1]) {
                KLP_CUDA(cudaFuncSetAttribute(k_rebin<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rebin_smem_bytes<T>(4096)));
                t->rebin_attr_set[sizeof(T) == 4 ? 0 : 1] = true;
            }
            k_rebin<T><<<(unsigned)nb, 256, rebin_smem_bytes<T>(r.n_stage), st>>>(r);
            KLP_CUDA(cudaGetLastError());
        }
        if (conv && !stop_after_profile) {
            TimedLaunch tl(t, st, 3);
            ConvArgs ca;
            ca.g = (const double*)t->conv_grid.p;
            ca.bw = (const double*)t->conv_grid.p + kConvN;
            ca.ybw = (const double*)t->conv_grid.p + 2 * kConvN;
            ca.prof = (const double*)t->prof.p;
            ca.prof_stride = kConvN - 1;
            ca.x = d_energy;
            ca.avg = (const double*)t->conv_avg.p;
            ca.spec = per_set ? d_spec + (size_t)off * n_flux : d_spec;
            ca.per_set = per_set;
            ca.n = n_flux;
            ca.out = d_out + (size_t)off * n_flux;
            ca.regular = &cc_ptr<T>(t)->edges_regular;
            dim3 grid((unsigned)((n_flux + kConvThreads - 1) / kConvThreads), (unsigned)nb);
            // gridDim.y limit is 65535
            if (nb > 65535) return fail(KLP_ERR_ARG, "chunk must be <= 65535 for convolution models");
            // CTAs per set: one when the batch fills the GPU (the per-set table is then built once), more for small batches
            const long long fill = std::max<long long>(1, (4LL * t->sm_count + nb - 1) / nb);
            if (t->conv_mode == 1) {          // cumulative-profile formulation
                constexpr int EPL = 2;
                const int chunks = (n_flux + 32 * EPL - 2) / (32 * EPL - 1);
                const unsigned cps = (unsigned)std::max<long long>(1, std::min<long long>((chunks + 7) / 8, fill));
                if (t->conv_src_smem && n_flux <= 4096) {   // {avg, spectrum} pairs staged in shared memory too
                    const size_t smem = kConvCumSrcOffset + sizeof(double2) * (size_t)n_flux;
                    const unsigned cps2 = (unsigned)std::max<long long>(1, std::min<long long>((chunks + 15) / 16, fill));
                    KLP_CUDA(cudaFuncSetAttribute(k_conv_cum<EPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kConvCumSrcOffset + sizeof(double2) * 4096)));
                    k_conv_cum<EPL, true><<<dim3(cps2, (unsigned)nb), 512, smem, st>>>(ca);
                } else {
                    KLP_CUDA(cudaFuncSetAttribute(k_conv_cum<EPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvFastSmemBytes));
                    k_conv_cum<EPL, false><<<dim3(cps, (unsigned)nb), 256, kConvFastSmemBytes, st>>>(ca);
                }
            } else if (t->conv_mode == 0) {   // the reference's operation order (default)
                const unsigned cps = (unsigned)std::max<long long>(1, std::min<long long>((n_flux + kConvXThreads - 1) / kConvXThreads, fill));
                KLP_CUDA(cudaFuncSetAttribute(k_conv_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvXSmemBytes));
                k_conv_exact<<<dim3(cps, (unsigned)nb), kConvXThreads, kConvXSmemBytes, st>>>(ca);
            } else if (t->conv_mode == 3) {   // round-1 kernels, kept for A/B measurements
                dim3 gridf((unsigned)((n_flux + kConvFastBinsPerCta - 1) / kConvFastBinsPerCta), (unsigned)nb);
                KLP_CUDA(cudaFuncSetAttribute(k_conv_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvFastSmemBytes));
                k_conv_fast<<<gridf, kConvFastThreads, kConvFastSmemBytes, st>>>(ca);
            } else {
                KLP_CUDA(cudaFuncSetAttribute(k_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmemBytes));
                k_conv<<<grid, kConvThreads, kConvSmemBytes, st>>>(ca);
            }
            KLP_CUDA(cudaGetLastError());
            // caller edges that are not strictly ascending and positive: the reference's own loops (returns at once otherwise)
            k_conv_general<<<dim3((unsigned)std::min<long long>((n_flux + 255) / 256, fill), (unsigned)nb), 256, 0, st>>>(ca);
            KLP_CUDA(cudaGetLastError());
        }
    }
    return KLP_OK;
}

int eval_device(klp_table* t, int model, int precision, long long B, const double* d_params, const double* d_energy,
                int n_flux, const double* d_spec, int per_set, double* d_out, cudaStream_t st, const EvalExtra& ex,
                bool stop_after_profile = false) {
    if (!t) return fail(KLP_ERR_ARG, "null table");
    if (model_nparams(model) < 0) return fail(KLP_ERR_ARG, "unknown model");
    if (precision != KLP_F32 && precision != KLP_F64) return fail(KLP_ERR_ARG, "unknown precision");
    if (B < 0 || n_flux < 1) return fail(KLP_ERR_ARG, "bad batch or grid size");
    if (B == 0) return KLP_OK;
    if (!d_params || !d_energy || (!d_out && !stop_after_profile)) return fail(KLP_ERR_ARG, "null buffer");
    if (model >= 2 && !d_spec && !stop_after_profile) return fail(KLP_ERR_ARG, "convolution models need an input spectrum");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);   // host-side bookkeeping only; the device work is ordered by the caller's stream
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(KLP_ERR_CUDA, "no usable CUDA device for this table (host-only table or device lost); there is no CPU path");
    if (!ex.setup_valid) t->setup_precision = -1;   // k_setup is about to overwrite the per-call constants
    if (precision == KLP_F32)
        return eval_device_t<float>(t, model, B, d_params, d_energy, n_flux, d_spec, per_set, d_out, st, ex, stop_after_profile);
    return eval_device_t<double>(t, model, B, d_params, d_energy, n_flux, d_spec, per_set, d_out, st, ex, stop_after_profile);
}

int ensure_streams(klp_table* t) {
    if (t->s_compute) return KLP_OK;
    KLP_CUDA(cudaStreamCreateWithFlags(&t->s_compute, cudaStreamNonBlocking));
    KLP_CUDA(cudaStreamCreateWithFlags(&t->s_h2d, cudaStreamNonBlocking));
    KLP_CUDA(cudaStreamCreateWithFlags(&t->s_d2h, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        KLP_CUDA(cudaEventCreateWithFlags(&t->ev_h2d[k], cudaEventDisableTiming));
        KLP_CUDA(cudaEventCreateWithFlags(&t->ev_comp[k], cudaEventDisableTiming));
        KLP_CUDA(cudaEventCreateWithFlags(&t->ev_d2h[k], cudaEventDisableTiming));
    }
    return KLP_OK;
}

// Host-buffer batch: chunked, double-buffered through pinned memory; H2D, compute and D2H run on
// three streams so copies overlap the kernels.
// Staging chunks of the host-buffer pipeline for B sets of n_flux bins: full slots of stage_mb megabytes of output (at least 16
// and at most 65535 sets: very wide grids get fewer sets per slot, not more pinned memory), then a SHORT final chunk.  The
// copy-out of the last chunk (D2H + the copy from the pinned slot into the caller's buffer: ~15 ms for a 128 MB slot) overlaps
// nothing, so it is cut to stage_tail sets, at most a quarter slot (default 4096: ~4 ms exposed, and its ~30 ms of kernels cover
// the copy-out of the chunk before it; measured on 1e5 kline5 sets: 128 MB slots 133.0 k spectra/s, 256 MB slots without the
// short chunk 127.9 k, with it 134.5 k).  Host logic only (klp_stage_chunk_plan exposes it to the CPU tests).
std::vector<long long> stage_chunks(long long B, int n_flux, long long stage_mb, long long stage_tail) {
    long long hc = (stage_mb << 20) / ((long long)n_flux * 8);
    hc = std::max<long long>(16, std::min<long long>(hc, 65535));
    hc = std::min(hc, B);
    long long tail = std::min<long long>(stage_tail, hc / 4);
    if (tail < 1 || B <= hc + tail) tail = 0;
    std::vector<long long> cuts;
    long long off = 0;
    const long long body = B - tail;
    while (off < body) { const long long nb = std::min(hc, body - off); cuts.push_back(nb); off += nb; }
    if (tail > 0) cuts.push_back(tail);
    return cuts;
}

int eval_host(klp_table* t, int model, int precision, long long B, const double* params, const double* energy,
              int n_flux, const double* spectrum, int per_set, double* out, const EvalExtra& ex, double* lims_out) {
    if (!t) return fail(KLP_ERR_ARG, "null table");
    const int P = model_nparams(model);
    if (P < 0) return fail(KLP_ERR_ARG, "unknown model");
    if (B < 0 || n_flux < 1) return fail(KLP_ERR_ARG, "bad batch or grid size");
    if (B == 0) return KLP_OK;
    if (!params || !energy || !out) return fail(KLP_ERR_ARG, "null buffer");
    const bool conv = model >= 2;
    if (conv && !spectrum) return fail(KLP_ERR_ARG, "convolution models need an input spectrum");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);
    NvtxRange nvtx_call("klp_eval_batch");
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(KLP_ERR_CUDA, "no usable CUDA device for this table (host-only table or device lost); there is no CPU path");
    int rc;
    if ((rc = ensure_streams(t))) return rc;
    // XSPEC and samplers call again and again with the same energy grid: when the per-call constants on the device
    // (k_setup: fine grid, g* knots, convolution grid and 2/(x_i + x_{i+1})) are those of this grid, precision and model
    // kind, neither the grid is copied nor k_setup launched again.
    const bool setup_cached = t->setup_precision == precision && t->setup_conv == (conv ? 1 : 0) &&
                              t->setup_energy.size() == (size_t)n_flux + 1 &&
                              std::memcmp(t->setup_energy.data(), energy, sizeof(double) * ((size_t)n_flux + 1)) == 0;
    if ((rc = t->d_energy.ensure(sizeof(double) * (size_t)(n_flux + 1)))) return rc;
    if (!setup_cached)
        KLP_CUDA(cudaMemcpyAsync(t->d_energy.p, energy, sizeof(double) * (size_t)(n_flux + 1), cudaMemcpyHostToDevice, t->s_compute));
    if (conv && !per_set) {
        if ((rc = t->d_spec.ensure(sizeof(double) * (size_t)n_flux))) return rc;
        KLP_CUDA(cudaMemcpyAsync(t->d_spec.p, spectrum, sizeof(double) * (size_t)n_flux, cudaMemcpyHostToDevice, t->s_compute));
    }
    // (no synchronisation needed here: copies from pageable memory are staged before cudaMemcpyAsync returns, and the kernels
    // that read the grid / spectrum are enqueued on the same stream)
    // staging chunk: t->stage_mb of output per slot (two slots; default 256 MB).  Larger chunks amortise the idle tail of
    // every k_quad launch, smaller ones shorten the last, un-overlapped copy-out.
    const std::vector<long long> cuts = stage_chunks(B, n_flux, t->stage_mb, t->stage_tail);
    const long long hc = *std::max_element(cuts.begin(), cuts.end());   // sets per staging slot
    const size_t out_bytes = sizeof(double) * (size_t)hc * n_flux + 16;   // (+16: the radial-grid end points of a legacy call ride behind the spectra)
    const size_t par_bytes = sizeof(double) * (size_t)hc * P;
    for (int k = 0; k < 2; ++k) {
        if ((rc = t->d_params[k].ensure(par_bytes)) || (rc = t->d_out[k].ensure(out_bytes)) ||
            (rc = t->h_params[k].ensure(par_bytes)) || (rc = t->h_out[k].ensure(out_bytes)))
            return rc;
        if (conv && per_set)
            if ((rc = t->d_spec_slot[k].ensure(out_bytes)) || (rc = t->h_spec_slot[k].ensure(out_bytes))) return rc;
    }
    const long long n_chunks = (long long)cuts.size();
    if ((rc = t->h_lims.ensure(16))) return rc;
    if (n_chunks == 1) {
        // One staging chunk (the XSPEC call pattern is B = 1): everything on ONE stream, one synchronisation -- no events, no
        // cross-stream waits; the radial-grid end points of the ratchet ride along with the spectra.
        cudaStream_t s = t->s_compute;
        std::memcpy(t->h_params[0].p, params, sizeof(double) * (size_t)B * P);
        KLP_CUDA(cudaMemcpyAsync(t->d_params[0].p, t->h_params[0].p, sizeof(double) * (size_t)B * P, cudaMemcpyHostToDevice, s));
        const double* dspec = (const double*)t->d_spec.p;
        if (conv && per_set) {
            std::memcpy(t->h_spec_slot[0].p, spectrum, sizeof(double) * (size_t)B * n_flux);
            KLP_CUDA(cudaMemcpyAsync(t->d_spec_slot[0].p, t->h_spec_slot[0].p, sizeof(double) * (size_t)B * n_flux, cudaMemcpyHostToDevice, s));
            dspec = (const double*)t->d_spec_slot[0].p;
        }
        EvalExtra e2 = ex;
        e2.setup_valid = setup_cached;
        const bool lims = lims_out && ex.want_lims;
        const size_t out_len = sizeof(double) * (size_t)B * n_flux;
        if (lims) e2.lims_dst = (char*)t->d_out[0].p + out_len;   // 16 spare bytes behind the spectra (see the buffer sizes above)
        rc = eval_device(t, model, precision, B, (const double*)t->d_params[0].p, (const double*)t->d_energy.p, n_flux, dspec, per_set,
                         (double*)t->d_out[0].p, s, e2);
        if (rc) return rc;
        KLP_CUDA(cudaMemcpyAsync(t->h_out[0].p, t->d_out[0].p, out_len + (lims ? 16 : 0), cudaMemcpyDeviceToHost, s));
        KLP_CUDA(cudaStreamSynchronize(s));
        std::memcpy(out, t->h_out[0].p, out_len);
        if (lims) {
            const char* lp = (const char*)t->h_out[0].p + out_len;
            if (precision == KLP_F32) { lims_out[0] = ((const float*)lp)[0]; lims_out[1] = ((const float*)lp)[1]; }
            else { lims_out[0] = ((const double*)lp)[0]; lims_out[1] = ((const double*)lp)[1]; }
        }
        if (!setup_cached) {   // remember what the per-call constants on the device belong to
            t->setup_energy.assign(energy, energy + n_flux + 1);
            t->setup_precision = precision;
            t->setup_conv = conv ? 1 : 0;
        }
        return KLP_OK;
    }
    long long pending_off[2] = {-1, -1}, pending_n[2] = {0, 0};
    auto drain = [&](int k) -> int {
        if (pending_off[k] < 0) return KLP_OK;
        KLP_CUDA(cudaEventSynchronize(t->ev_d2h[k]));
        std::memcpy(out + (size_t)pending_off[k] * n_flux, t->h_out[k].p, sizeof(double) * (size_t)pending_n[k] * n_flux);
        pending_off[k] = -1;
        return KLP_OK;
    };
    long long off = 0;
    for (long long c = 0; c < n_chunks; off += cuts[(size_t)c], ++c) {
        const int k = (int)(c & 1);
        const long long nb = cuts[(size_t)c];
        if ((rc = drain(k))) return rc;
        std::memcpy(t->h_params[k].p, params + (size_t)off * P, sizeof(double) * (size_t)nb * P);
        KLP_CUDA(cudaMemcpyAsync(t->d_params[k].p, t->h_params[k].p, sizeof(double) * (size_t)nb * P, cudaMemcpyHostToDevice, t->s_h2d));
        const double* dspec = (const double*)t->d_spec.p;
        if (conv && per_set) {
            std::memcpy(t->h_spec_slot[k].p, spectrum + (size_t)off * n_flux, sizeof(double) * (size_t)nb * n_flux);
            KLP_CUDA(cudaMemcpyAsync(t->d_spec_slot[k].p, t->h_spec_slot[k].p, sizeof(double) * (size_t)nb * n_flux, cudaMemcpyHostToDevice, t->s_h2d));
            dspec = (const double*)t->d_spec_slot[k].p;
        }
        KLP_CUDA(cudaEventRecord(t->ev_h2d[k], t->s_h2d));
        KLP_CUDA(cudaStreamWaitEvent(t->s_compute, t->ev_h2d[k], 0));
        EvalExtra e2 = ex;
        e2.want_lims = ex.want_lims && c == 0;
        e2.setup_valid = setup_cached || c > 0;
        rc = eval_device(t, model, precision, nb, (const double*)t->d_params[k].p, (const double*)t->d_energy.p, n_flux,
                         dspec, per_set, (double*)t->d_out[k].p, t->s_compute, e2);
        if (rc) return rc;
        KLP_CUDA(cudaEventRecord(t->ev_comp[k], t->s_compute));
        KLP_CUDA(cudaStreamWaitEvent(t->s_d2h, t->ev_comp[k], 0));
        KLP_CUDA(cudaMemcpyAsync(t->h_out[k].p, t->d_out[k].p, sizeof(double) * (size_t)nb * n_flux, cudaMemcpyDeviceToHost, t->s_d2h));
        KLP_CUDA(cudaEventRecord(t->ev_d2h[k], t->s_d2h));
        pending_off[k] = off;
        pending_n[k] = nb;
    }
    if ((rc = drain(0)) || (rc = drain(1))) return rc;
    if (lims_out && ex.want_lims) {
        if (precision == KLP_F32) {
            float l[2];
            KLP_CUDA(cudaMemcpy(l, t->lims.p, sizeof(l), cudaMemcpyDeviceToHost));
            lims_out[0] = l[0];
            lims_out[1] = l[1];
        } else {
            KLP_CUDA(cudaMemcpy(lims_out, t->lims.p, sizeof(double) * 2, cudaMemcpyDeviceToHost));
        }
    }
    KLP_CUDA(cudaGetLastError());
    if (!setup_cached) {   // remember what the per-call constants on the device belong to
        t->setup_energy.assign(energy, energy + n_flux + 1);
        t->setup_precision = precision;
        t->setup_conv = conv ? 1 : 0;
    }
    return KLP_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* klp_last_error(void) { return g_err.c_str(); }

int klp_model_nparams(int model) { return model_nparams(model); }

int klp_device_count(int* n) {
    if (!n) return fail(KLP_ERR_ARG, "null argument");
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; return fail(KLP_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    return KLP_OK;
}

int klp_table_create(const float* spins, int n_a, const float* mus, int n_mu, int n_r, int n_g, const float* radii,
                     const float* gmin, const float* gmax, const float* upper, const float* lower, int device,
                     klp_table** out) {
    if (!out) return fail(KLP_ERR_ARG, "null out");
    *out = nullptr;
    if (!spins || !mus || !radii || !gmin || !gmax || !upper || !lower) return fail(KLP_ERR_ARG, "null table array");
    if (n_a < 1 || n_mu < 1 || n_r < 3 || n_g < 2 || n_g > kMaxNG) return fail(KLP_ERR_ARG, "bad table dimensions");
    for (int i = 1; i < n_a; ++i) if (!(spins[i] > spins[i - 1])) return fail(KLP_ERR_ARG, "spins must be strictly ascending");
    for (int i = 1; i < n_mu; ++i) if (!(mus[i] > mus[i - 1])) return fail(KLP_ERR_ARG, "mus must be strictly ascending");
    const size_t frame_len = (size_t)n_r * (2 * n_g + 3);
    if (device == -1) {
        // host-only table: index logic works, every evaluation entry point refuses (no CPU path)
        klp_table* t = new klp_table;
        t->device = -1;
        t->n_a = n_a; t->n_mu = n_mu; t->n_r = n_r; t->n_g = n_g;
        t->spins.assign(spins, spins + n_a);
        t->mus.assign(mus, mus + n_mu);
        t->frame_stride = (frame_len + 3) & ~(size_t)3;
        const size_t F = (size_t)n_a * n_mu;
        t->h_frames.assign(F * t->frame_stride, 0.f);
        for (size_t f = 0; f < F; ++f) {
            float* dst = t->h_frames.data() + f * t->frame_stride;
            std::memcpy(dst, radii + f * n_r, sizeof(float) * n_r);
            std::memcpy(dst + n_r, gmin + f * n_r, sizeof(float) * n_r);
            std::memcpy(dst + 2 * (size_t)n_r, gmax + f * n_r, sizeof(float) * n_r);
            std::memcpy(dst + 3 * (size_t)n_r, upper + f * (size_t)n_r * n_g, sizeof(float) * (size_t)n_r * n_g);
            std::memcpy(dst + 3 * (size_t)n_r + (size_t)n_r * n_g, lower + f * (size_t)n_r * n_g, sizeof(float) * (size_t)n_r * n_g);
        }
        *out = t;
        return KLP_OK;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(KLP_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(KLP_ERR_ARG, "bad device ordinal");
    DeviceGuard dg(device);
    if (!dg.ok) return fail(KLP_ERR_CUDA, "cannot select CUDA device");
    klp_table* t = new klp_table;
    t->device = device;
    t->n_a = n_a; t->n_mu = n_mu; t->n_r = n_r; t->n_g = n_g;
    t->spins.assign(spins, spins + n_a);
    t->mus.assign(mus, mus + n_mu);
    t->frame_stride = (frame_len + 3) & ~(size_t)3;
    const size_t F = (size_t)n_a * n_mu;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete t; return fail(KLP_ERR_CUDA, "cudaGetDeviceProperties"); }
    t->sm_count = prop.multiProcessorCount;
    t->coop_supported = prop.cooperativeLaunch != 0;
    t->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    t->max_smem_sm = (int)prop.sharedMemPerMultiprocessor;
    auto bail = [&](int code, const std::string& m) { klp_table_destroy(t); return fail(code, m); };
    if (cudaMalloc(&t->d_frames, sizeof(float) * F * t->frame_stride) != cudaSuccess) return bail(KLP_ERR_NOMEM, "cudaMalloc(frames)");
    if (cudaMalloc(&t->d_spins, sizeof(float) * n_a) != cudaSuccess) return bail(KLP_ERR_NOMEM, "cudaMalloc(spins)");
    if (cudaMalloc(&t->d_mus, sizeof(float) * n_mu) != cudaSuccess) return bail(KLP_ERR_NOMEM, "cudaMalloc(mus)");
    {
        // operand-range precondition of the FAST arithmetic, per frame: branches exactly 0 or in [2^-20, 2^20] (the reference
        // zero-fills radii a frame does not cover; a zero numerator is exact in every FAST sequence, and what blending a zero
        // with its neighbours produces is checked per set on the blended frame, k_quad phase 1), g limits in [2^-10, 2^10].
        // A set may use the FAST sequences when its four frames pass.
        std::vector<unsigned char> okv(F, 1);
        for (size_t f = 0; f < F; ++f) {
            bool ok = true;
            const float* u = upper + f * (size_t)n_r * n_g;
            const float* l = lower + f * (size_t)n_r * n_g;
            for (size_t i = 0; i < (size_t)n_r * n_g && ok; ++i)
                ok = (u[i] == 0.f || (u[i] >= 9.5367431640625e-07f && u[i] <= 1048576.f)) && (l[i] == 0.f || (l[i] >= 9.5367431640625e-07f && l[i] <= 1048576.f));
            const float* gn = gmin + f * n_r;
            const float* gx = gmax + f * n_r;
            for (int i = 0; i < n_r && ok; ++i)
                ok = gn[i] >= 0.0009765625f && gn[i] <= 1024.f && gx[i] >= 0.0009765625f && gx[i] <= 1024.f;
            okv[f] = ok ? 1 : 0;
            t->frames_fast += ok ? 1 : 0;
        }
        if (cudaMalloc(&t->d_frame_ok, F) != cudaSuccess) return bail(KLP_ERR_NOMEM, "cudaMalloc(frame flags)");
        if (cudaMemcpy(t->d_frame_ok, okv.data(), F, cudaMemcpyHostToDevice) != cudaSuccess) return bail(KLP_ERR_CUDA, "cudaMemcpy(frame flags)");
    }
    // pack frames: radii | gmin | gmax | upper | lower, padded to the stride; staged in slabs
    const size_t slab_frames = std::max<size_t>(1, (64u << 20) / (sizeof(float) * t->frame_stride));
    std::vector<float> slab(slab_frames * t->frame_stride, 0.f);
    for (size_t f0 = 0; f0 < F; f0 += slab_frames) {
        const size_t nf = std::min(slab_frames, F - f0);
        for (size_t k = 0; k < nf; ++k) {
            const size_t f = f0 + k;
            float* dst = slab.data() + k * t->frame_stride;
            std::memcpy(dst, radii + f * n_r, sizeof(float) * n_r);
            std::memcpy(dst + n_r, gmin + f * n_r, sizeof(float) * n_r);
            std::memcpy(dst + 2 * (size_t)n_r, gmax + f * n_r, sizeof(float) * n_r);
            std::memcpy(dst + 3 * (size_t)n_r, upper + f * (size_t)n_r * n_g, sizeof(float) * (size_t)n_r * n_g);
            std::memcpy(dst + 3 * (size_t)n_r + (size_t)n_r * n_g, lower + f * (size_t)n_r * n_g, sizeof(float) * (size_t)n_r * n_g);
        }
        if (cudaMemcpy(t->d_frames + f0 * t->frame_stride, slab.data(), sizeof(float) * nf * t->frame_stride,
                       cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(KLP_ERR_CUDA, "cudaMemcpy(frames)");
    }
    if (cudaMemcpy(t->d_spins, spins, sizeof(float) * n_a, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t->d_mus, mus, sizeof(float) * n_mu, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(KLP_ERR_CUDA, "cudaMemcpy(parameters)");
    t->store = std::make_shared<FrameStore>();
    t->store->device = device;
    t->store->d_frames = t->d_frames;
    t->store->d_spins = t->d_spins;
    t->store->d_mus = t->d_mus;
    t->store->d_frame_ok = t->d_frame_ok;
    *out = t;
    return KLP_OK;
}

int klp_table_share(klp_table* t, klp_table** out) {
    if (!out) return fail(KLP_ERR_ARG, "null out");
    *out = nullptr;
    if (!t) return fail(KLP_ERR_ARG, "null table");
    if (t->device < 0 || !t->store) return fail(KLP_ERR_CUDA, "only device-resident tables can be shared (there is no CPU path)");
    klp_table* w = new klp_table;
    w->device = t->device;
    w->store = t->store;
    w->n_a = t->n_a; w->n_mu = t->n_mu; w->n_r = t->n_r; w->n_g = t->n_g;
    w->frame_stride = t->frame_stride;
    w->spins = t->spins;
    w->mus = t->mus;
    w->d_frames = t->d_frames; w->d_spins = t->d_spins; w->d_mus = t->d_mus; w->d_frame_ok = t->d_frame_ok;
    w->frames_fast = t->frames_fast;
    w->sm_count = t->sm_count; w->max_smem_optin = t->max_smem_optin; w->max_smem_sm = t->max_smem_sm;
    w->coop_supported = t->coop_supported;
    *out = w;
    return KLP_OK;
}

// ---- FITS reader for the reference's table schema (io.zig:8-87; cfitsio is not available, so this is
// a dependency-free reader of exactly what readFitsFile touches): HDU 2 column 1 = spins, HDU 3
// column 1 = cos(inclination), HDU 4.. = one BINTABLE per (a, mu) frame with columns
// 1 r, 2 gmin, 3 gmax, 4 upper_f[n_g], 5 lower_f[n_g] (HDU and column numbers 1-based as in cfitsio).
namespace {
struct FitsCol { char type; long repeat; size_t offset, width; };
struct FitsHdu {
    bool bintable = false;
    long naxis1 = 0, naxis2 = 0;
    size_t data_off = 0, data_bytes = 0;
    std::vector<FitsCol> cols;
};

bool fits_parse(FILE* f, std::vector<FitsHdu>& hdus, std::string& err) {
    std::fseek(f, 0, SEEK_END);
    const long fsize = std::ftell(f);
    long pos = 0;
    char card[81];
    card[80] = 0;
    while (pos + 2880 <= fsize) {
        FitsHdu h;
        long bitpix = 8, naxis = 0, pcount = 0, gcount = 1, tfields = 0;
        std::vector<long> naxes;
        std::vector<std::string> tform;
        bool end = false, first = true;
        while (!end) {
            if (pos + 2880 > fsize) { err = "truncated FITS header"; return false; }
            std::fseek(f, pos, SEEK_SET);
            for (int c = 0; c < 36; ++c) {
                if (std::fread(card, 1, 80, f) != 80) { err = "short read in FITS header"; return false; }
                std::string key(card, 8);
                while (!key.empty() && key.back() == ' ') key.pop_back();
                std::string val(card + 10, 70);
                if (first) {
                    first = false;
                    if (hdus.empty()) { if (key != "SIMPLE") { err = "not a FITS file"; return false; } }
                    else if (key != "XTENSION") { err = "bad FITS extension header"; return false; }
                    if (key == "XTENSION") h.bintable = val.find("BINTABLE") != std::string::npos;
                }
                else if (!end && (key == "XTENSION" || key == "SIMPLE")) { err = "FITS header without END card (the next HDU starts inside it)"; return false; }
                if (key == "END") { end = true; continue; }
                if (end || card[8] != '=') continue;
                auto num = [&]() { return std::strtol(val.c_str(), nullptr, 10); };
                auto str = [&]() {
                    size_t a = val.find('\''), b = val.find('\'', a + 1);
                    std::string v = (a == std::string::npos || b == std::string::npos) ? std::string() : val.substr(a + 1, b - a - 1);
                    while (!v.empty() && v.back() == ' ') v.pop_back();
                    return v;
                };
                if (key == "BITPIX") bitpix = num();
                else if (key == "NAXIS") { naxis = num(); naxes.assign((size_t)std::max(0l, naxis), 0); }
                else if (key == "PCOUNT") pcount = num();
                else if (key == "GCOUNT") gcount = num();
                else if (key == "TFIELDS") { tfields = num(); tform.assign((size_t)std::max(0l, tfields), ""); }
                else if (key.compare(0, 5, "NAXIS") == 0 && key.size() > 5) {
                    long k = std::strtol(key.c_str() + 5, nullptr, 10);
                    if (k >= 1 && k <= (long)naxes.size()) naxes[k - 1] = num();
                } else if (key.compare(0, 5, "TFORM") == 0) {
                    long k = std::strtol(key.c_str() + 5, nullptr, 10);
                    if (k >= 1 && k <= (long)tform.size()) tform[k - 1] = str();
                }
            }
            pos += 2880;
        }
        size_t nelem = naxes.empty() ? 0 : 1;
        for (long n : naxes) nelem *= (size_t)n;
        h.data_off = (size_t)pos;
        h.data_bytes = (size_t)(std::labs(bitpix) / 8) * (size_t)gcount * ((size_t)pcount + nelem);
        if (h.bintable) {
            if (naxes.size() != 2) { err = "BINTABLE without NAXIS=2"; return false; }
            h.naxis1 = naxes[0];
            h.naxis2 = naxes[1];
            size_t off = 0;
            for (auto& tf : tform) {
                char* endp = nullptr;
                long rep = std::strtol(tf.c_str(), &endp, 10);
                if (endp == tf.c_str()) rep = 1;
                const char ty = *endp;
                size_t w = ty == 'E' || ty == 'J' ? 4 : ty == 'D' || ty == 'K' ? 8 : ty == 'I' ? 2 : ty == 'B' || ty == 'L' || ty == 'A' ? 1 : 0;
                if (!w) { err = "unsupported TFORM '" + tf + "'"; return false; }
                h.cols.push_back({ty, rep, off, w});
                off += w * (size_t)rep;
            }
            if ((long)off != h.naxis1) { err = "TFORM widths do not add up to NAXIS1"; return false; }
        }
        pos += (long)((h.data_bytes + 2879) / 2880 * 2880);
        hdus.push_back(h);
    }
    return true;
}

// column `col` (0-based) of a BINTABLE as floats, rows x repeat
bool fits_column(FILE* f, const FitsHdu& h, size_t col, std::vector<float>& out, long& repeat, std::string& err) {
    if (!h.bintable || col >= h.cols.size()) { err = "missing table column"; return false; }
    const FitsCol& c = h.cols[col];
    if (c.type != 'E' && c.type != 'D') { err = "table column is not floating point"; return false; }
    repeat = c.repeat;
    std::vector<unsigned char> raw((size_t)h.naxis1 * (size_t)h.naxis2);
    std::fseek(f, (long)h.data_off, SEEK_SET);
    if (std::fread(raw.data(), 1, raw.size(), f) != raw.size()) { err = "truncated FITS table data"; return false; }
    out.resize((size_t)h.naxis2 * (size_t)c.repeat);
    for (long r = 0; r < h.naxis2; ++r)
        for (long k = 0; k < c.repeat; ++k) {
            const unsigned char* p = raw.data() + (size_t)r * h.naxis1 + c.offset + (size_t)k * c.width;
            if (c.type == 'E') {
                uint32_t u = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
                float v;
                std::memcpy(&v, &u, 4);
                out[(size_t)r * c.repeat + k] = v;
            } else {
                uint64_t u = 0;
                for (int b = 0; b < 8; ++b) u = (u << 8) | p[b];
                double v;
                std::memcpy(&v, &u, 8);
                out[(size_t)r * c.repeat + k] = (float)v;
            }
        }
    return true;
}

int load_fits_table(const char* path, int device, klp_table** out) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(KLP_ERR_NOTFOUND, std::string("cannot open table file: ") + path);
    std::vector<FitsHdu> hdus;
    std::string err;
    auto bad = [&](const std::string& m) { std::fclose(f); return fail(KLP_ERR_IO, std::string(path) + ": " + m); };
    if (!fits_parse(f, hdus, err)) return bad(err);
    if (hdus.size() < 4) return bad("expected spins (HDU 2), inclinations (HDU 3) and at least one frame (HDU 4..)");
    std::vector<float> spins, mus, col;
    long rep = 0;
    if (!fits_column(f, hdus[1], 0, spins, rep, err) || rep != 1) return bad("HDU 2: " + err);
    if (!fits_column(f, hdus[2], 0, mus, rep, err) || rep != 1) return bad("HDU 3: " + err);
    const size_t F = hdus.size() - 3;
    if (F != spins.size() * mus.size()) return bad("number of frame HDUs != n_spins * n_inclinations");
    const long n_r = hdus[3].naxis2;
    long n_g = 0;
    std::vector<float> radii, gmin, gmax, upper, lower;
    for (size_t fr = 0; fr < F; ++fr) {
        const FitsHdu& h = hdus[3 + fr];
        if (h.naxis2 != n_r) return bad("frames have different numbers of radii");
        std::vector<float>* dst3[3] = {&radii, &gmin, &gmax};
        for (int c = 0; c < 3; ++c) {
            if (!fits_column(f, h, (size_t)c, col, rep, err) || rep != 1) return bad("frame column: " + err);
            dst3[c]->insert(dst3[c]->end(), col.begin(), col.end());
        }
        std::vector<float>* dst2[2] = {&upper, &lower};
        for (int c = 0; c < 2; ++c) {
            if (!fits_column(f, h, (size_t)(3 + c), col, rep, err)) return bad("frame branch column: " + err);
            if (fr == 0 && c == 0) n_g = rep;
            if (rep != n_g) return bad("frames have different numbers of g* knots");
            dst2[c]->insert(dst2[c]->end(), col.begin(), col.end());
        }
    }
    std::fclose(f);
    return klp_table_create(spins.data(), (int)spins.size(), mus.data(), (int)mus.size(), (int)n_r, (int)n_g, radii.data(),
                            gmin.data(), gmax.data(), upper.data(), lower.data(), device, out);
}
}  // namespace

int klp_table_load(const char* path, int device, klp_table** out) {
    if (!out) return fail(KLP_ERR_ARG, "null out");
    *out = nullptr;
    if (!path) return fail(KLP_ERR_ARG, "null path");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(KLP_ERR_NOTFOUND, std::string("cannot open table file: ") + path);
    char magic[8];
    int32_t d[4];
    if (std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "SIMPLE  ", 8) == 0) {
        std::fclose(f);
        return load_fits_table(path, device, out);
    }
    std::fseek(f, 0, SEEK_SET);
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "KLPT0001", 8) != 0 || std::fread(d, 4, 4, f) != 4) {
        std::fclose(f);
        return fail(KLP_ERR_IO, std::string("neither a FITS nor a KLPT0001 table: ") + path);
    }
    const int n_a = d[0], n_mu = d[1], n_r = d[2], n_g = d[3];
    if (n_a < 1 || n_mu < 1 || n_r < 3 || n_g < 2 || n_g > kMaxNG) { std::fclose(f); return fail(KLP_ERR_IO, "bad table dimensions in file"); }
    const size_t F = (size_t)n_a * n_mu, per = (size_t)n_r * (2 * n_g + 3);
    std::vector<float> spins(n_a), mus(n_mu), body(F * per);
    bool ok = std::fread(spins.data(), 4, n_a, f) == (size_t)n_a && std::fread(mus.data(), 4, n_mu, f) == (size_t)n_mu &&
              std::fread(body.data(), 4, F * per, f) == F * per;
    std::fclose(f);
    if (!ok) return fail(KLP_ERR_IO, std::string("truncated table file: ") + path);
    std::vector<float> radii(F * n_r), gmin(F * n_r), gmax(F * n_r), upper(F * (size_t)n_r * n_g), lower(F * (size_t)n_r * n_g);
    for (size_t fr = 0; fr < F; ++fr) {
        const float* src = body.data() + fr * per;
        std::memcpy(&radii[fr * n_r], src, 4 * (size_t)n_r);
        std::memcpy(&gmin[fr * n_r], src + n_r, 4 * (size_t)n_r);
        std::memcpy(&gmax[fr * n_r], src + 2 * (size_t)n_r, 4 * (size_t)n_r);
        std::memcpy(&upper[fr * (size_t)n_r * n_g], src + 3 * (size_t)n_r, 4 * (size_t)n_r * n_g);
        std::memcpy(&lower[fr * (size_t)n_r * n_g], src + 3 * (size_t)n_r + (size_t)n_r * n_g, 4 * (size_t)n_r * n_g);
    }
    return klp_table_create(spins.data(), n_a, mus.data(), n_mu, n_r, n_g, radii.data(), gmin.data(), gmax.data(),
                            upper.data(), lower.data(), device, out);
}

void klp_table_destroy(klp_table* t) {
    if (!t) return;
    {
        // a table that is still installed for the XSPEC symbols must not dangle: uninstall it first
        std::lock_guard<std::recursive_mutex> lk(g_legacy_mu);
        if (g_default_table == t) g_default_table = nullptr;
        if (g_owned_table == t) g_owned_table = nullptr;
    }
    if (t->device >= 0) {
        DeviceGuard dg(t->device);
        cudaDeviceSynchronize();
        if (t->store) {
            t->store.reset();   // the frames go with the last table / workspace that references them
        } else {                // (a table whose construction failed half way)
            if (t->d_frames) cudaFree(t->d_frames);
            if (t->d_spins) cudaFree(t->d_spins);
            if (t->d_mus) cudaFree(t->d_mus);
            if (t->d_frame_ok) cudaFree(t->d_frame_ok);
        }
        for (Buf* b : {&t->cc32, &t->cc64, &t->fine, &t->prof, &t->conv_grid, &t->conv_avg, &t->scratch, &t->counter, &t->lims, &t->vals, &t->done, &t->order, &t->phase,
                       &t->d_energy, &t->d_spec, &t->d_params[0], &t->d_params[1], &t->d_out[0], &t->d_out[1],
                       &t->d_spec_slot[0], &t->d_spec_slot[1]})
            b->release();
        for (PinBuf* b : {&t->h_params[0], &t->h_params[1], &t->h_out[0], &t->h_out[1], &t->h_spec_slot[0], &t->h_spec_slot[1], &t->h_lims})
            b->release();
        if (t->s_compute) cudaStreamDestroy(t->s_compute);
        if (t->s_h2d) cudaStreamDestroy(t->s_h2d);
        if (t->s_d2h) cudaStreamDestroy(t->s_d2h);
        for (int k = 0; k < 2; ++k) {
            if (t->ev_h2d[k]) cudaEventDestroy(t->ev_h2d[k]);
            if (t->ev_comp[k]) cudaEventDestroy(t->ev_comp[k]);
            if (t->ev_d2h[k]) cudaEventDestroy(t->ev_d2h[k]);
        }
        for (auto& p : t->ev_pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    }
    delete t;
}

int klp_table_dims(const klp_table* t, int* dims) {
    if (!t || !dims) return fail(KLP_ERR_ARG, "null argument");
    dims[0] = t->n_a; dims[1] = t->n_mu; dims[2] = t->n_r; dims[3] = t->n_g;
    return KLP_OK;
}

int klp_table_parameters(const klp_table* t, float* spins, float* mus) {
    if (!t) return fail(KLP_ERR_ARG, "null table");
    if (spins) std::memcpy(spins, t->spins.data(), sizeof(float) * t->spins.size());
    if (mus) std::memcpy(mus, t->mus.data(), sizeof(float) * t->mus.size());
    return KLP_OK;
}

int klp_table_read_frame(const klp_table* t, size_t frame, float* out) {
    if (!t || !out) return fail(KLP_ERR_ARG, "null argument");
    const size_t F = (size_t)t->n_a * t->n_mu, len = (size_t)t->n_r * (2 * t->n_g + 3);
    if (frame >= F) return fail(KLP_ERR_ARG, "frame index out of range");
    if (t->device < 0) {
        std::memcpy(out, t->h_frames.data() + frame * t->frame_stride, sizeof(float) * len);
        return KLP_OK;
    }
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(KLP_ERR_CUDA, "cannot select the table's device");
    KLP_CUDA(cudaMemcpy(out, t->d_frames + frame * t->frame_stride, sizeof(float) * len, cudaMemcpyDeviceToHost));
    return KLP_OK;
}

// line-profile.zig:60-70
size_t klp_table_index(const klp_table* t, size_t i_a, size_t i_mu) { return i_a * (size_t)t->n_mu + i_mu; }

// line-profile.zig:72-93 (first >=, none -> 0)
void klp_find_parameter_indices(const klp_table* t, float a, float mu, size_t* out2) {
    out2[0] = 0;
    out2[1] = 0;
    for (size_t i = 0; i < t->spins.size(); ++i) if (t->spins[i] >= a) { out2[0] = i; break; }
    for (size_t i = 0; i < t->mus.size(); ++i) if (t->mus[i] >= mu) { out2[1] = i; break; }
}

// line-profile.zig:209-231
void klp_find_tables(const klp_table* t, const size_t* pi, size_t* out4) {
    size_t p0 = pi[0], p1 = pi[1];
    out4[1] = klp_table_index(t, p0, p1);
    if (p0 > 0) p0 -= 1;
    out4[0] = klp_table_index(t, p0, p1);
    if (p1 > 0) p1 -= 1;
    out4[2] = klp_table_index(t, p0, p1);
    if (p0 < pi[0]) p0 += 1;
    out4[3] = klp_table_index(t, p0, p1);
}

// line-profile.zig:196-207
int klp_parameter_weight(const klp_table* t, int which, float value, size_t index, float* weight) {
    if (index == 0) return 0;
    const std::vector<float>& arr = which == 0 ? t->spins : t->mus;
    if (index >= arr.size()) return 0;
    *weight = (value - arr[index - 1]) / (arr[index] - arr[index - 1]);
    return 1;
}

int klp_eval_batch_device(klp_table* t, int model, int precision, long B, const double* d_params, const double* d_energy,
                          int n_flux, const double* d_spectrum, int spectrum_per_set, double* d_out, void* cuda_stream) {
    EvalExtra ex;
    return eval_device(t, model, precision, B, d_params, d_energy, n_flux, d_spectrum, spectrum_per_set, d_out,
                       (cudaStream_t)cuda_stream, ex);
}

int klp_eval_batch(klp_table* t, int model, int precision, long B, const double* params, const double* energy, int n_flux,
                   const double* spectrum, int spectrum_per_set, double* out) {
    EvalExtra ex;
    return eval_host(t, model, precision, B, params, energy, n_flux, spectrum, spectrum_per_set, out, ex, nullptr);
}

int klp_eval_debug(klp_table* t, int model, int precision, long B, const double* params, const double* energy, int n_flux,
                   double* fine_out, double* profile_out) {
    if (!t) return fail(KLP_ERR_ARG, "null table");
    const int P = model_nparams(model);
    if (P < 0 || B < 1 || n_flux < 1 || !params || !energy) return fail(KLP_ERR_ARG, "bad argument");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(KLP_ERR_CUDA, "no usable CUDA device for this table (host-only table or device lost); there is no CPU path");
    if (B > t->chunk) return fail(KLP_ERR_ARG, "klp_eval_debug: B must fit one chunk");
    int rc;
    if ((rc = ensure_streams(t))) return rc;
    Buf dpar, den, dout;
    auto cleanup = [&]() { dpar.release(); den.release(); dout.release(); };
    if ((rc = dpar.ensure(sizeof(double) * (size_t)B * P)) || (rc = den.ensure(sizeof(double) * (size_t)(n_flux + 1))) ||
        (rc = dout.ensure(sizeof(double) * (size_t)B * n_flux))) { cleanup(); return rc; }
    if (cudaMemcpy(dpar.p, params, sizeof(double) * (size_t)B * P, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(den.p, energy, sizeof(double) * (size_t)(n_flux + 1), cudaMemcpyHostToDevice) != cudaSuccess) {
        cleanup();
        return fail(KLP_ERR_CUDA, "klp_eval_debug: host-to-device copy failed");
    }
    EvalExtra ex;
    rc = eval_device(t, model, precision, B, (const double*)dpar.p, (const double*)den.p, n_flux, nullptr, 0, (double*)dout.p,
                     t->s_compute, ex, /*stop_after_profile=*/true);
    if (rc) { cleanup(); return rc; }
    cudaError_t e = cudaStreamSynchronize(t->s_compute);
    if (e != cudaSuccess) { cleanup(); return fail(KLP_ERR_CUDA, std::string("kernel failed: ") + cudaGetErrorString(e)); }
    if (fine_out) {
        if (precision == KLP_F32) {
            std::vector<float> h((size_t)B * kFineG);
            if (cudaMemcpy(h.data(), t->fine.p, sizeof(float) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) { cleanup(); return fail(KLP_ERR_CUDA, "klp_eval_debug: device-to-host copy failed"); }
            for (long s = 0; s < B; ++s)
                for (int i = 0; i < kNFine; ++i) fine_out[(size_t)s * kNFine + i] = h[(size_t)s * kFineG + i];
        } else {
            std::vector<double> h((size_t)B * kFineG);
            if (cudaMemcpy(h.data(), t->fine.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) { cleanup(); return fail(KLP_ERR_CUDA, "klp_eval_debug: device-to-host copy failed"); }
            for (long s = 0; s < B; ++s)
                for (int i = 0; i < kNFine; ++i) fine_out[(size_t)s * kNFine + i] = h[(size_t)s * kFineG + i];
        }
    }
    if (profile_out && model >= 2 &&
        cudaMemcpy(profile_out, t->prof.p, sizeof(double) * (size_t)B * (kConvN - 1), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cleanup();
        return fail(KLP_ERR_CUDA, "klp_eval_debug: device-to-host copy failed");
    }
    cleanup();
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(KLP_ERR_CUDA, cudaGetErrorString(e));
    return KLP_OK;
}

int klp_blend_frames_device(klp_table* t, int precision, long B, const double* d_a_incl, void* d_out, void* cuda_stream) {
    if (!t || !d_a_incl || !d_out || B < 0) return fail(KLP_ERR_ARG, "bad argument");
    if (B == 0) return KLP_OK;
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(KLP_ERR_CUDA, "no usable CUDA device for this table (host-only table or device lost); there is no CPU path");
    const int grid = (int)std::min<long long>(B, (long long)t->sm_count * 8);
    if (precision == KLP_F32) {
        BlendArgs<float> a{table_dev(t), d_a_incl, B, (float*)d_out};
        k_blend<float><<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(a);
    } else if (precision == KLP_F64) {
        BlendArgs<double> a{table_dev(t), d_a_incl, B, (double*)d_out};
        k_blend<double><<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(a);
    } else {
        return fail(KLP_ERR_ARG, "unknown precision");
    }
    KLP_CUDA(cudaGetLastError());
    return KLP_OK;
}

int klp_set_timing(klp_table* t, int enable) {
    if (!t) return fail(KLP_ERR_ARG, "null table");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);
    t->timing = enable != 0;
    t->ev_next = 0;
    t->ev_used.clear();
    return KLP_OK;
}

int klp_get_timing(klp_table* t, double* ms4, long* launches4) {
    if (!t) return fail(KLP_ERR_ARG, "null table");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);
    DeviceGuard dg(t->device);
    for (int k = 0; k < 4; ++k) { t->ms[k] = 0; t->launches[k] = 0; }
    for (auto& u : t->ev_used) {
        auto& p = t->ev_pool[u.second];
        KLP_CUDA(cudaEventSynchronize(p.second));
        float ms = 0;
        KLP_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
        t->ms[u.first] += ms;
        t->launches[u.first] += 1;
    }
    for (int k = 0; k < 4; ++k) {
        if (ms4) ms4[k] = t->ms[k];
        if (launches4) launches4[k] = t->launches[k];
    }
    t->ev_next = 0;   // the next call starts a new accumulation period
    t->ev_used.clear();
    return KLP_OK;
}

int klp_last_launch_info(klp_table* t, char* buf, size_t len) {
    if (!t || !buf || len == 0) return fail(KLP_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);
    std::snprintf(buf, len, "%s fast_frames=%zu/%zu", t->quad_variant, t->frames_fast, (size_t)t->n_a * t->n_mu);
    return KLP_OK;
}

int klp_get_phase_times(klp_table* t, double* us9) {
    if (!t || !us9) return fail(KLP_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);
    DeviceGuard dg(t->device);
    if (!dg.ok || !t->phase.p) return fail(KLP_ERR_ARG, "no phase stamps: set option phase_clock and evaluate first");
    unsigned long long h[16];
    KLP_CUDA(cudaMemcpy(h, t->phase.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 10; ++k) us9[k] = h[k] ? (double)(h[k] - h[0]) * 1e-3 : -1.0;
    return KLP_OK;
}

// The staging chunks klp_eval_batch would use (sizes[cap]); returns their number, -1 on bad arguments.  No device needed.
int klp_stage_chunk_plan(long B, int n_flux, long stage_mb, long stage_tail, long* sizes, int cap) {
    if (B < 1 || n_flux < 1 || stage_mb < 1 || stage_tail < 0 || !sizes || cap < 1) return -1;
    const std::vector<long long> cuts = stage_chunks(B, n_flux, stage_mb, stage_tail);
    if ((int)cuts.size() > cap) return -1;
    for (size_t i = 0; i < cuts.size(); ++i) sizes[i] = (long)cuts[i];
    return (int)cuts.size();
}

int klp_set_option(klp_table* t, const char* key, long value) {
    if (!t || !key) return fail(KLP_ERR_ARG, "null argument");
    std::lock_guard<std::recursive_mutex> lk(t->mu_);
    std::string k(key);
    if (k == "quad_threads") {
        if (value != 0 && value != 256 && value != 512 && value != 768 && value != 1024) return fail(KLP_ERR_ARG, "quad_threads");
        t->quad_threads = (int)value;
    } else if (k == "splits") {
        if (value < 0 || value > kMaxSplits) return fail(KLP_ERR_ARG, "splits");
        t->splits = (int)value;
    } else if (k == "order") {
        t->use_order = value != 0;
    } else if (k == "stage_mb") {
        if (value < 1 || value > 4096) return fail(KLP_ERR_ARG, "stage_mb");
        t->stage_mb = (int)value;
    } else if (k == "phase_clock") {
        t->phase_clock = value != 0;
    } else if (k == "coop") {
        t->use_coop = value != 0;
    } else if (k == "deposit_ring") {
        t->deposit_ring = value != 0;
    } else if (k == "ring_discard") {
        t->ring_discard = value != 0;
    } else if (k == "frame_global") {
        t->frame_global = value != 0;
    } else if (k == "stage_tail") {
        if (value < 0 || value > 65535) return fail(KLP_ERR_ARG, "stage_tail");
        t->stage_tail = (int)value;
    } else if (k == "conv_src_smem") {
        t->conv_src_smem = value ? 1 : 0;
    } else if (k == "conv_mode") {
        if (value < 0 || value > 3) return fail(KLP_ERR_ARG, "conv_mode");
        t->conv_mode = (int)value;
    } else if (k == "fast") {
        t->use_fast = value != 0;
    } else if (k == "chunk") {
        if (value < 1 || value > (1 << 20)) return fail(KLP_ERR_ARG, "chunk");
        t->chunk = value;
    } else {
        return fail(KLP_ERR_ARG, "unknown option");
    }
    return KLP_OK;
}

int klp_arith_selftest(klp_table* t, unsigned long long seed, unsigned long long n_total, unsigned long long* mismatches4) {
    if (!t || !mismatches4) return fail(KLP_ERR_ARG, "null argument");
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(KLP_ERR_CUDA, "no usable CUDA device for this table");
    Buf d;
    int rc;
    if ((rc = d.ensure(4 * sizeof(unsigned long long)))) return rc;
    cudaMemset(d.p, 0, 4 * sizeof(unsigned long long));
    const int grid = t->sm_count * 8, threads = grid * 256;
    const unsigned long long per = (n_total + threads - 1) / threads;
    k_arith_selftest<<<grid, 256>>>(seed, per, (unsigned long long*)d.p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(mismatches4, d.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    d.release();
    if (e != cudaSuccess) return fail(KLP_ERR_CUDA, cudaGetErrorString(e));
    return KLP_OK;
}

int klp_fma_peak_gflops(klp_table* t, int precision, double* gflops) {
    if (!t || !gflops) return fail(KLP_ERR_ARG, "null argument");
    DeviceGuard dg(t->device);
    Buf sink;
    int rc;
    if ((rc = sink.ensure(64))) return rc;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = t->sm_count * 8, iters = 1 << 15;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        if (precision == KLP_F32) k_fma_peak<float><<<grid, 256>>>((float*)sink.p, iters);
        else k_fma_peak<double><<<grid, 256>>>((double*)sink.p, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double fl = (double)grid * 256.0 * iters * 8.0 * 2.0;
        if (ms > 0) best = std::max(best, fl / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    sink.release();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(KLP_ERR_CUDA, cudaGetErrorString(e));
    *gflops = best;
    return KLP_OK;
}

// ---- XSPEC model surface (xspec-wrapper.zig:425-530) ----------------------------------------
namespace {
const char* kMsgPrefix = "kerrlineprofile: ";              // xspec-wrapper.zig:12
const char* kDataDirEnv = "KLINE_PROF_DATA_DIR";           // xspec-wrapper.zig:10
const char* kModelFile = "kerr-transfer-functions.fits";   // MODELPATH, xspec-wrapper.zig:11
const char* kModelFileFlat = "kerr-transfer-functions.klpt"; // flat-container twin (lineprofiles_b200/tableio.py)
int g_legacy_precision = KLP_F32;
bool g_have_lims = false;              // Q2 ratchet state
double g_lims[2] = {0, 0};

int legacy_table(klp_table** out) {
    if (g_default_table) { *out = g_default_table; return KLP_OK; }
    if (!g_owned_table) {
        const char* root = std::getenv(kDataDirEnv);
        std::string path = std::string(root ? root : ".") + "/" + kModelFile;  // getModelPath, :30-48
        int dev = 0;
        if (const char* d = std::getenv("KLINE_PROF_DEVICE")) dev = std::atoi(d);
        if (FILE* probe = std::fopen(path.c_str(), "rb")) std::fclose(probe);
        else path = std::string(root ? root : ".") + "/" + kModelFileFlat;
        int rc = klp_table_load(path.c_str(), dev, &g_owned_table);
        if (rc) return rc;
    }
    *out = g_owned_table;
    return KLP_OK;
}

void legacy_call(int model, const double* energy, int n_flux, const double* params, double* flux) {
    std::lock_guard<std::recursive_mutex> lk(g_legacy_mu);
    klp_table* t = nullptr;
    int rc = legacy_table(&t);
    if (rc == KLP_OK) {
        EvalExtra ex;
        ex.use_limits = g_have_lims;
        ex.lim_lo = g_lims[0];
        ex.lim_hi = g_lims[1];
        ex.want_lims = true;
        std::vector<double> spec;
        const double* sp = nullptr;
        if (model >= 2) { spec.assign(flux, flux + n_flux); sp = spec.data(); }  // flux_copy, :200-201
        std::vector<double> outv((size_t)std::max(n_flux, 1));
        double lims[2];
        rc = eval_host(t, model, g_legacy_precision, 1, params, energy, n_flux, sp, 0, outv.data(), ex, lims);
        if (rc == KLP_OK) {
            std::memcpy(flux, outv.data(), sizeof(double) * (size_t)n_flux);
            g_have_lims = true;
            g_lims[0] = lims[0];
            g_lims[1] = lims[1];
            return;
        }
    }
    // xspec-wrapper.zig:361-372
    if (rc == KLP_ERR_NOTFOUND)   // only a missing file gets the reference's hint; a present but unreadable one reports the parser's message
        std::fprintf(stderr, "error: %sFailed to find table model. Set the %s environment variable to where %s is located\n",
                     kMsgPrefix, kDataDirEnv, kModelFile);
    else
        std::fprintf(stderr, "error: %s%s\n", kMsgPrefix, g_err.c_str());
    std::fprintf(stderr, "error: %sMODEL OUTPUT SET TO ZERO\n", kMsgPrefix);
    if (flux && n_flux > 0) std::memset(flux, 0, sizeof(double) * (size_t)n_flux);
}
}  // namespace

void klp_set_default_table(klp_table* t) {
    std::lock_guard<std::recursive_mutex> lk(g_legacy_mu);
    g_default_table = t;
}
void klp_set_legacy_precision(int precision) {
    std::lock_guard<std::recursive_mutex> lk(g_legacy_mu);
    if (precision == KLP_F32 || precision == KLP_F64) g_legacy_precision = precision;
    g_have_lims = false;
}
void klp_legacy_reset(void) {
    std::lock_guard<std::recursive_mutex> lk(g_legacy_mu);
    g_have_lims = false;
}

void kline(const double* energy, int n_flux, const double* params, int, double* flux, double*, const char*) {
    legacy_call(KLP_KLINE, energy, n_flux, params, flux);
}
void kline5(const double* energy, int n_flux, const double* params, int, double* flux, double*, const char*) {
    legacy_call(KLP_KLINE5, energy, n_flux, params, flux);
}
void kconv(const double* energy, int n_flux, const double* params, int, double* flux, double*, const char*) {
    legacy_call(KLP_KCONV, energy, n_flux, params, flux);
}
void kconv5(const double* energy, int n_flux, const double* params, int, double* flux, double*, const char*) {
    legacy_call(KLP_KCONV5, energy, n_flux, params, flux);
}
void kconv10(const double* energy, int n_flux, const double* params, int, double* flux, double*, const char*) {
    legacy_call(KLP_KCONV10, energy, n_flux, params, flux);
}

}  // extern "C"

#include "klp_group.inc"
