// C-ABI layer: context, pattern-side precompute and the batched scan pipeline.
// See include/apd_b200.h for the contract and the reference interfaces each call replaces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/apd_b200.h"
#include "internal.h"
#include "loudness.h"
#include "peaks.h"

using namespace apd;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(APD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)

namespace {

struct ClipHost {
    int L = 0, sw = 0, group = 0, strategy = 0, is_short = 0;
    double tone_hz = 0.0, lufs = 0.0;     // tone_hz: NaN = the clip takes the normal verifier
    float self_max = 0.0f;
    float* d_raw = nullptr;
    float* d_norm = nullptr;
    float* d_rev = nullptr;
    float2* d_spec = nullptr;       // group-size spectrum of the reversed normalised clip
    float* d_self_corr = nullptr;   // 2L-1
    float* d_win_cache = nullptr;
    int tone_P = 0;
    double2* d_tone_chirp = nullptr;
    double2* d_tone_tw = nullptr;
    double2* d_tone_pre = nullptr;
    double2* d_tone_post = nullptr;
};

struct Group {
    int sw = 0, halo = 0;
    std::vector<int> clips;
    int shape = 0;                  // index into apd_ctx::shapes
    long long spec_off = 0;
    int max_L = 0;
};

// Clips whose groups share one four-step FFT shape are correlated by the same launches.
struct ShapeClass {
    Fft4Plan plan;
    std::vector<int> clips;
    int* d_clips = nullptr;
    std::vector<int> groups;              // sliding-window groups of this shape
    int* d_group_halo = nullptr;
    int* d_group_index = nullptr;
    long long* d_group_spec_off = nullptr;
};

template <typename T>
cudaError_t dalloc(T** p, size_t n)
{
    return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
}

template <typename T>
cudaError_t upload(T** p, const std::vector<T>& v)
{
    cudaError_t e = dalloc(p, v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

__global__ void k_prepare_clip(const float* __restrict__ raw, int L, const double* __restrict__ gain,
                               float* __restrict__ norm, float* __restrict__ rev)
{
    const double g = *gain;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        // reference lib.rs:220-227 (clip side keeps NaN; a silent clip is not a valid pattern)
        double v = (double)raw[i] * g;
        if (v == v) v = v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v);
        const float f = (float)v;
        norm[i] = f;
        rev[L - 1 - i] = f;
    }
}

// clip-side window-max down-sampling (reference lib.rs:283-318), one block per window
__global__ void k_window_max(const float* __restrict__ cc, int lo, int hi, int ds, float* __restrict__ out)
{
    const int nw = hi - lo;
    const double step = (double)nw / (double)ds;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ds; i += gridDim.x * blockDim.x) {
        int a = (int)((double)i * step);
        int b = (int)((double)(i + 1) * step);
        if (b <= a) b = a + 1;
        if (a >= nw) a = nw - 1;
        if (b > nw) b = nw;
        float m = -INFINITY;
        for (int t = a; t < b; ++t) m = fmaxf(m, cc[lo + t]);
        out[i] = m;
    }
}

// Ordered compaction (one CTA) of the units of one shape class whose normalised maximum reaches the
// height threshold: only those can have a peak (every other sample of the unit is smaller).
// Appends to sel[] starting at *begin (end of the previous shape's units) and writes the new end.
__global__ void __launch_bounds__(1024)
k_select_shape(const unsigned int* __restrict__ unit_max_bits, const float* __restrict__ self_max,
               const int* __restrict__ shape_clips, int ns, int n_clips, int n_chunks, float height,
               int2* __restrict__ sel, const int* __restrict__ begin, int* __restrict__ end, int capacity,
               int* __restrict__ overflow)
{
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = *begin;
    __syncthreads();
    const int n_units = n_chunks * ns;
    for (int u0 = 0; u0 < n_units; u0 += blockDim.x) {
        const int u = u0 + threadIdx.x;
        bool pass = false;
        int ci = 0, clip = 0;
        if (u < n_units) {
            ci = u / ns;
            clip = shape_clips[u % ns];
            const float am = __uint_as_float(unit_max_bits[(long long)ci * n_clips + clip]);
            const float mc = fmaxf(self_max[clip], am);
            pass = mc > 0.0f && (am / mc) >= height;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int i = 0; i < w; ++i) off += warp_tot[i];
        if (pass) {
            const int pos = off + __popc(bal & ((1u << lane) - 1));
            if (pos < capacity) sel[pos] = make_int2(ci, clip);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += warp_tot[i];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (base > capacity) { atomicOr(overflow, 8); base = capacity; }
        *end = base;
    }
}

__global__ void k_record_npeaks(const int2* __restrict__ sel, const int* __restrict__ sel_count, int slot0, int nslots,
                                const int* __restrict__ n_peaks, int n_clips, int* __restrict__ unit_npeaks)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots || slot0 + s >= *sel_count) return;
    const int2 u = sel[slot0 + s];
    unit_npeaks[(long long)u.x * n_clips + u.y] = n_peaks[s];
}

}  // namespace

struct apd_ctx {
    int device = 0, sr = 0, n_clips = 0, maxB = 0;
    long long C = 0;
    float height = kDefaultHeight;
    std::vector<ClipHost> clips;
    std::vector<Group> groups;
    std::vector<ShapeClass> shapes;
    int max_halo = 0;
    KwConfig kw{};
    std::map<int, Fft4Plan> self_plans;
    long long launches = 0;

    // per-clip device tables
    int *d_clip_len = nullptr, *d_clip_group = nullptr, *d_strategy = nullptr, *d_is_short = nullptr;
    int *d_win_lo = nullptr, *d_win_hi = nullptr, *d_win_ds = nullptr, *d_tone_P = nullptr;
    float* d_self_max = nullptr;
    const float** d_self_corr_ptrs = nullptr;
    const float** d_win_cache_ptrs = nullptr;
    const float2** d_clip_spec_ptrs = nullptr;
    double *d_tone_hz = nullptr, *d_tone_thr = nullptr;
    const double2** d_tone_chirp_ptrs = nullptr;
    const double2** d_tone_tw_ptrs = nullptr;
    const double2** d_tone_pre_ptrs = nullptr;
    const double2** d_tone_post_ptrs = nullptr;
    long long* d_clip_spec_off = nullptr;
    int2* d_sel = nullptr;                // selected (ci, clip) units of the batch, shape-class major
    int sel_capacity = 0;
    SectionGeom* d_geoms = nullptr;
    SectionGeom* h_geoms = nullptr;       // pinned

    // workspace
    int cells_stride = 0;
    double *d_kw_state = nullptr, *d_kw_energy = nullptr, *d_kw_em1 = nullptr, *d_kw_patch = nullptr;
    double *d_lufs = nullptr, *d_gain = nullptr;
    long long spec_slab = 0;              // complex elements per chunk (sum of group M)
    float2* d_spec = nullptr;
    long long scratch_elems = 0;
    float2* d_scratch = nullptr;
    void* d_unit_desc = nullptr;          // per-unit descriptors of one inverse launch (corr_inv.cu)
    int inv_units = 0;                    // units per inverse launch
    unsigned int* d_unit_max = nullptr;
    int* d_unit_npeaks = nullptr;
    // device counters: [0..S] begin of each shape class in d_sel ([S] = total selected),
    //                  [S+1] candidates emitted, [S+2] overflow flags, [S+3] tone work items
    int* d_counts = nullptr;
    int* h_counts = nullptr;              // pinned mirror
    int n_slots = 0;
    long long corr_stride = 0, cand_stride = 0;
    int peak_stride = 0;
    float* d_corr = nullptr;
    int* d_cand_idx = nullptr;
    float* d_cand_val = nullptr;
    unsigned char* d_cand_state = nullptr;
    int* d_peaks = nullptr;
    float* d_peak_height = nullptr;
    int *d_n_peaks = nullptr, *d_n_cands = nullptr;
    apd_candidate* d_slot_cands = nullptr;
    apd_candidate* d_out = nullptr;
    int out_capacity = 0;
    // tone
    int tone_ctas = 0, tone_wl = 0, tone_item_cap = 0;      // tone_ctas: work items per tone round
    unsigned char* d_tone_alive = nullptr;                  // per work item: flanks still needed
    bool tone_all_segments = false;                         // a trace / single-candidate call wants every metric
    int tone_max_P = 0, tone_max_L = 0;
    double* d_tone_stats = nullptr;
    long long tone_stride = 0;
    double2* d_tone_scratch = nullptr;
    void* d_tone_items = nullptr;
    double* d_tone_metrics = nullptr;

    // phase 2 (selected units only) runs on a side stream, overlapped with phase 1 of the next sub-batch:
    // it has its own four-step scratch / descriptors, and everything phase 1 hands over to it is
    // double-buffered (BatchSet below; the pointers above are those of the set in use).
    float2* d_scratch2 = nullptr;
    void* d_unit_desc2 = nullptr;
    cudaStream_t side = nullptr;          // phase 2
    cudaStream_t pre = nullptr;           // loudness of the next sub-batch (latency-bound, hidden under phase 1)
    cudaStream_t corr2 = nullptr;         // odd correlate launches: the row pass (HBM-bound) of one launch overlaps the
                                          // column pass (latency-bound) of the previous one
    cudaEvent_t corr_fork = nullptr, corr_join = nullptr;
    float2* d_scratch_b = nullptr;        // second four-step intermediate / descriptor set for that stream
    void* d_unit_desc_b = nullptr;
    cudaStream_t fwd = nullptr;           // opt-in (APD_B200_FWD_STREAM=1): forward FFT of the next sub-batch, own scratch, beside
                                          // the correlate stage of the current one
    float2* d_scratch_fwd = nullptr;      // four-step intermediate of the forward FFT (= d_scratch unless that stream is on)
    cudaStream_t tone = nullptr;          // deferred marker-tone verification: bandwidth-heavy f64 FFT passes, at the
                                          // caller's (low) priority so that they fill the tails of the correlate launches
    cudaEvent_t scan_start = nullptr, tone_go = nullptr, tone_done = nullptr;
    struct BatchSet {
        float2* d_spec = nullptr;
        unsigned int* d_unit_max = nullptr;
        int* d_unit_npeaks = nullptr;
        int *d_counts = nullptr, *h_counts = nullptr;
        int2* d_sel = nullptr;
        double *d_lufs = nullptr, *d_gain = nullptr;
        SectionGeom *d_geoms = nullptr, *h_geoms = nullptr;
        // results of the sub-batch: per set, so that the marker-tone verification of one sub-batch (tone stream)
        // can still be appending records while the side stream already verifies the next sub-batch
        apd_candidate* d_out = nullptr;
        apd_candidate* h_out = nullptr;   // pinned staging of d_out
        void* d_tone_items = nullptr;
        cudaEvent_t out_ready = nullptr;  // records and counters of the sub-batch are in h_out / h_counts
        int n_expected = 0;
        int chunk_begin = 0, chunk_end = 0;
        cudaEvent_t p1_done = nullptr, loud_done = nullptr;
        cudaEvent_t fwd_done = nullptr;
        cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    } sets[3];
    int n_sets = 2;                       // sets in use; APD_B200_SETS=3: the tone verification of a sub-batch also overlaps the
                                          // normal phase 2 of the next one (measured: same step time)
    int cur_set = 0;

    // optional stage timing (CUDA events: [0..1] loudness stream, [2..3] forward FFT, [7]..[4] correlate stage on the
    // caller's stream, [5..6] phase 2)
    bool profile = false;
    cudaEvent_t* ev = nullptr;
    double stage_ms[4] = {0, 0, 0, 0};
    long long work[4] = {0, 0, 0, 0};     // selected units, candidate records, tone work items, sub-batches

    // state of the staged batch
    const float* audio = nullptr;
    long long base = 0, nsamp = 0;
    int chunk_begin = 0, chunk_end = 0;
    bool staged = false;
};

// the sets in use (the third exists only with APD_B200_SETS=3)
struct UsedSets {
    apd_ctx::BatchSet *b, *e;
    apd_ctx::BatchSet* begin() const { return b; }
    apd_ctx::BatchSet* end() const { return e; }
};
static UsedSets used_sets(apd_ctx* c) { return UsedSets{c->sets, c->sets + c->n_sets}; }

static void use_set(apd_ctx* c, int k)
{
    apd_ctx::BatchSet& cur = c->sets[c->cur_set];
    cur.chunk_begin = c->chunk_begin;
    cur.chunk_end = c->chunk_end;
    apd_ctx::BatchSet& b = c->sets[k];
    c->d_spec = b.d_spec; c->d_unit_max = b.d_unit_max; c->d_unit_npeaks = b.d_unit_npeaks;
    c->d_counts = b.d_counts; c->h_counts = b.h_counts; c->d_sel = b.d_sel;
    c->d_lufs = b.d_lufs; c->d_gain = b.d_gain; c->d_geoms = b.d_geoms; c->h_geoms = b.h_geoms;
    c->d_out = b.d_out; c->d_tone_items = b.d_tone_items;
    c->chunk_begin = b.chunk_begin; c->chunk_end = b.chunk_end;
    c->ev = b.ev;
    c->cur_set = k;
}

static void fill_geoms(apd_ctx* c)
{
    for (size_t g = 0; g < c->groups.size(); ++g) {
        SectionGeom G;
        G.audio = c->audio;
        G.base = c->base;
        G.total = c->base + c->nsamp;
        G.chunk = c->C;
        G.chunk0 = c->chunk_begin;
        G.halo = c->groups[g].halo;
        c->h_geoms[g] = G;
    }
}

static SectionGeom union_geom(const apd_ctx* c)
{
    SectionGeom G = c->h_geoms[0];
    G.halo = c->max_halo;
    return G;
}

static UnitCtx unit_ctx(const apd_ctx* c)
{
    return UnitCtx{c->d_geoms, c->d_clip_group, c->d_clip_len, c->d_clip_spec_off, c->d_clip_spec_ptrs};
}

extern "C" const char* apd_last_error(void) { return g_err.c_str(); }

// |corr(clip, clip)| / max through the same kernels as the sections (reference apd.py:373-383).
struct InitTmp {
    unsigned int* max_bits = nullptr;     // [1]
    int* ints = nullptr;                  // [0] = 0 (clip index / group), [1] = clip length
    float* zero_f = nullptr;              // [1] = 0
    const float2** spec_ptr = nullptr;    // [1]
    long long* zero_ll = nullptr;         // [1] = 0
    SectionGeom* geom = nullptr;          // [1]
    void* desc = nullptr;                 // one unit descriptor
};

static int self_correlation(apd_ctx* c, ClipHost& cl, InitTmp& t)
{
    const int L = cl.L;
    const int M_min = L < 16 ? 16 : L;                       // real length 2M >= 2L-1
    std::string err;
    auto it = c->self_plans.find(M_min);
    if (it == c->self_plans.end()) {
        Fft4Plan P;
        if (!build_plan(M_min, &P, &err)) return fail(APD_ERR_UNSUPPORTED, err);
        it = c->self_plans.emplace(M_min, P).first;
    }
    const Fft4Plan& P = it->second;
    float2 *sa = nullptr, *sb = nullptr, *scr = nullptr;
    CK(dalloc(&sa, (size_t)P.M));
    CK(dalloc(&sb, (size_t)P.M));
    CK(dalloc(&scr, (size_t)P.M));
    SectionGeom Ga{cl.d_norm, 0, L, L, 0, 0}, Gb{cl.d_rev, 0, L, L, 0, 0};
    launch_forward(P, Ga, FwdGroups{nullptr, nullptr, nullptr, 0}, nullptr, 0, 1, scr, sa, P.M, 0);
    launch_forward(P, Gb, FwdGroups{nullptr, nullptr, nullptr, 0}, nullptr, 0, 1, scr, sb, P.M, 0);
    CK(cudaMemset(t.max_bits, 0, sizeof(unsigned int)));
    CK(cudaMemcpy(t.ints + 1, &L, sizeof(int), cudaMemcpyHostToDevice));
    const float2* hp = sb;
    CK(cudaMemcpy(t.spec_ptr, &hp, sizeof(hp), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t.geom, &Ga, sizeof(Ga), cudaMemcpyHostToDevice));
    UnitSrc U{nullptr, nullptr, t.ints, 1, 0, 0, 0, 0};                // dense: unit 0 = (section 0, clip 0)
    UnitCtx X{t.geom, t.ints, t.ints + 1, t.zero_ll, t.spec_ptr};
    InvOut O{t.max_bits, 1, cl.d_self_corr, (long long)(2 * L - 1), t.zero_f};
    launch_inverse(P, X, sa, P.M, U, 1, scr, t.desc, O, false, 0);
    launch_inverse(P, X, sa, P.M, U, 1, scr, t.desc, O, true, 0);
    CK(cudaDeviceSynchronize());
    unsigned int bits = 0;
    CK(cudaMemcpy(&bits, t.max_bits, sizeof(bits), cudaMemcpyDeviceToHost));
    memcpy(&cl.self_max, &bits, sizeof(float));
    c->launches += 8;
    cudaFree(sa);
    cudaFree(sb);
    cudaFree(scr);
    return APD_OK;
}

static int create_ctx(apd_ctx* c, int device, int sample_rate, int64_t chunk_samples, float height_min,
                          int n_clips, const apd_clip_desc* descs, int max_batch_chunks)
{
    CK(cudaSetDevice(device));
    c->device = device;
    c->sr = sample_rate;
    c->C = chunk_samples;
    c->n_clips = n_clips;
    c->maxB = max_batch_chunks;
    c->height = std::isnan(height_min) ? kDefaultHeight : height_min;      // NaN: not given (apd.py:520)
    std::string err;
    if (!kw_config_create(sample_rate, &c->kw, &err)) return fail(APD_ERR_UNSUPPORTED, err);

    // ---- clips and sliding-window groups (reference apd.py:155-161)
    c->clips.resize(n_clips);
    std::map<int, int> sw_to_group;
    for (int p = 0; p < n_clips; ++p) {
        ClipHost& cl = c->clips[p];
        if (!descs[p].samples || descs[p].length <= 0) return fail(APD_ERR_INVALID, "empty clip");
        cl.L = descs[p].length;
        cl.sw = (int)((cl.L + (long long)sample_rate - 1) / sample_rate);
        if ((long long)2 * cl.sw * sample_rate > chunk_samples) {
            return fail(APD_ERR_INVALID, "seconds_per_chunk is too small for clip " + std::to_string(p));
        }
        cl.strategy = descs[p].strategy;
        cl.tone_hz = descs[p].strategy == APD_STRATEGY_MARKER_TONE ? descs[p].tone_hz : nan("");   // NaN: no tone verifier
        cl.is_short = (2LL * cl.L < sample_rate) ? 1 : 0;             // L/sr < 0.5 (apd.py:628)
        auto it = sw_to_group.find(cl.sw);
        if (it == sw_to_group.end()) {
            it = sw_to_group.emplace(cl.sw, (int)c->groups.size()).first;
            Group g;
            g.sw = cl.sw;
            g.halo = cl.sw * sample_rate;
            c->groups.push_back(g);
        }
        cl.group = it->second;
        c->groups[cl.group].clips.push_back(p);
        c->groups[cl.group].max_L = std::max(c->groups[cl.group].max_L, cl.L);
    }
    const int G = (int)c->groups.size();
    long long max_nout = 0, max_M = 0;
    int min_L = c->clips[0].L;
    for (auto& cl : c->clips) min_L = std::min(min_L, cl.L);
    for (auto& g : c->groups) {
        const long long n_out = c->C + g.halo + g.max_L - 1;
        max_nout = std::max(max_nout, n_out);
        c->max_halo = std::max(c->max_halo, g.halo);
        Fft4Plan plan;
        if (plan_min_M_for(n_out) > (1LL << 21) || !build_plan((int)plan_min_M_for(n_out), &plan, &err)) {
            return fail(APD_ERR_UNSUPPORTED, err.empty() ? "chunk too long for the FFT plan" : err);
        }
        int shape = -1;
        for (size_t k = 0; k < c->shapes.size(); ++k)
            if (c->shapes[k].plan.N1 == plan.N1 && c->shapes[k].plan.N2 == plan.N2) shape = (int)k;
        if (shape < 0) {
            ShapeClass sc;
            sc.plan = plan;
            c->shapes.push_back(sc);
            shape = (int)c->shapes.size() - 1;
        } else {
            free_plan(&plan);
        }
        g.shape = shape;
        c->shapes[shape].groups.push_back((int)(&g - &c->groups[0]));
        for (int p : g.clips) c->shapes[shape].clips.push_back(p);
        g.spec_off = c->spec_slab;
        c->spec_slab += c->shapes[shape].plan.M;
        max_M = std::max<long long>(max_M, c->shapes[shape].plan.M);
    }
    size_t max_class_groups = 1;
    for (auto& sc : c->shapes) {
        std::sort(sc.clips.begin(), sc.clips.end());
        CK(upload(&sc.d_clips, sc.clips));
        std::vector<int> halo;
        std::vector<long long> off;
        for (int g : sc.groups) {
            halo.push_back(c->groups[g].halo);
            off.push_back(c->groups[g].spec_off);
        }
        CK(upload(&sc.d_group_halo, halo));
        CK(upload(&sc.d_group_index, sc.groups));
        CK(upload(&sc.d_group_spec_off, off));
        max_class_groups = std::max(max_class_groups, sc.groups.size());
    }
    const int S = (int)c->shapes.size();

    // ---- pattern-side precompute on the device
    InitTmp tmp;
    double *d_tl = nullptr, *d_tg = nullptr, *d_tstate = nullptr, *d_ten = nullptr, *d_tem1 = nullptr, *d_tpatch = nullptr;
    CK(dalloc(&tmp.max_bits, 1));
    CK(dalloc(&tmp.ints, 2));
    CK(cudaMemset(tmp.ints, 0, 2 * sizeof(int)));
    CK(dalloc(&tmp.zero_f, 1));
    CK(cudaMemset(tmp.zero_f, 0, sizeof(float)));
    CK(dalloc(&tmp.spec_ptr, 1));
    CK(dalloc(&tmp.zero_ll, 1));
    CK(cudaMemset(tmp.zero_ll, 0, sizeof(long long)));
    CK(dalloc(&tmp.geom, 1));
    CK(cudaMalloc(&tmp.desc, corr_inv_desc_bytes(1)));
    CK(dalloc(&d_tl, 1));
    CK(dalloc(&d_tg, 1));
    CK(dalloc(&d_tpatch, (size_t)c->kw.patch_cells));
    int max_L = 0;
    for (auto& cl : c->clips) max_L = std::max(max_L, cl.L);
    const int clip_cells = (c->kw.general ? 2 : 1) * ((max_L + c->kw.cell - 1) / c->kw.cell + 4);
    CK(dalloc(&d_tstate, (size_t)clip_cells * 4));
    CK(dalloc(&d_ten, (size_t)clip_cells));
    CK(dalloc(&d_tem1, 1));
    for (int p = 0; p < n_clips; ++p) {
        ClipHost& cl = c->clips[p];
        const int L = cl.L;
        CK(dalloc(&cl.d_raw, (size_t)L));
        CK(dalloc(&cl.d_norm, (size_t)L));
        CK(dalloc(&cl.d_rev, (size_t)L));
        CK(dalloc(&cl.d_self_corr, (size_t)2 * L - 1));
        CK(cudaMemcpy(cl.d_raw, descs[p].samples, sizeof(float) * L, cudaMemcpyHostToDevice));
        // loudness with block = clip_seconds if < 0.5 else 0.4 (apd.py:169-171): same rule as a section
        SectionGeom Gc{cl.d_raw, 0, L, L, 0, 0};
        CK(cudaMemcpy(tmp.geom, &Gc, sizeof(Gc), cudaMemcpyHostToDevice));
        launch_loudness(c->kw, Gc, &Gc, tmp.geom, 1, 1, clip_cells, d_tstate, d_ten, d_tem1, d_tpatch, d_tl, d_tg, 0);
        k_prepare_clip<<<std::min(256, (L + 255) / 256), 256>>>(cl.d_raw, L, d_tg, cl.d_norm, cl.d_rev);
        CK(cudaMemcpy(&cl.lufs, d_tl, sizeof(double), cudaMemcpyDeviceToHost));
        c->launches += 5;
        // spectrum of the reversed clip at the group's FFT size
        const Fft4Plan& gplan = c->shapes[c->groups[cl.group].shape].plan;
        CK(dalloc(&cl.d_spec, (size_t)gplan.M));
        {
            float2* scr = nullptr;
            CK(dalloc(&scr, (size_t)gplan.M));
            SectionGeom Gr{cl.d_rev, 0, L, L, 0, 0};
            launch_forward(gplan, Gr, FwdGroups{nullptr, nullptr, nullptr, 0}, nullptr, 0, 1, scr, cl.d_spec, gplan.M, 0);
            CK(cudaDeviceSynchronize());
            cudaFree(scr);
            c->launches += 2;
        }
        int rc = self_correlation(c, cl, tmp);
        if (rc != APD_OK) return rc;
    }

    // ---- per-clip tables
    std::vector<int> h_len(n_clips), h_group(n_clips), h_strat(n_clips), h_short(n_clips), h_lo(n_clips * 3),
        h_hi(n_clips * 3), h_ds(n_clips * 3, 0), h_P(n_clips, 0);
    std::vector<float> h_smax(n_clips);
    std::vector<const float*> h_cc(n_clips), h_cache(n_clips);
    std::vector<const float2*> h_spec(n_clips);
    std::vector<double> h_thz(n_clips), h_thr(n_clips * 6);
    std::vector<const double2*> h_chirp(n_clips, nullptr), h_tw(n_clips, nullptr), h_pre(n_clips, nullptr),
        h_post(n_clips, nullptr);
    int max_P = 0;
    for (int p = 0; p < n_clips; ++p) {
        ClipHost& cl = c->clips[p];
        const int L = cl.L, W = 2 * L - 1;
        h_len[p] = L; h_group[p] = cl.group; h_strat[p] = cl.strategy; h_short[p] = cl.is_short;
        h_smax[p] = cl.self_max; h_cc[p] = cl.d_self_corr; h_spec[p] = cl.d_spec; h_thz[p] = cl.tone_hz;
        // Pearson windows (apd.py:808-829); bounds use Python's round(): half to even
        const int wins[3][3] = {{0, 5, 252}, {4, 6, 101}, {5, 10, 252}};
        int total_ds = 0;
        if (cl.is_short) {
            h_lo[p * 3] = (int)std::nearbyint((double)(W * 0LL) / 10.0);
            h_hi[p * 3] = (int)std::nearbyint((double)((long long)W * 10) / 10.0);
            h_ds[p * 3] = 505;
            total_ds = 505;
        } else {
            for (int w = 0; w < 3; ++w) {
                h_lo[p * 3 + w] = (int)std::nearbyint((double)((long long)W * wins[w][0]) / 10.0);
                h_hi[p * 3 + w] = (int)std::nearbyint((double)((long long)W * wins[w][1]) / 10.0);
                h_ds[p * 3 + w] = wins[w][2];
                total_ds += wins[w][2];
            }
        }
        CK(dalloc(&cl.d_win_cache, (size_t)total_ds));
        int off = 0;
        for (int w = 0; w < 3; ++w) {
            if (h_ds[p * 3 + w] <= 0) continue;
            k_window_max<<<4, 128>>>(cl.d_self_corr, h_lo[p * 3 + w], h_hi[p * 3 + w], h_ds[p * 3 + w], cl.d_win_cache + off);
            off += h_ds[p * 3 + w];
            ++c->launches;
        }
        h_cache[p] = cl.d_win_cache;
        // marker-tone thresholds (apd.py:698-705) and Bluestein tables
        const double defaults[6] = {0.95, 0.80, 9.0, 0.92, 0.25, 0.65};
        const double given[6] = {descs[p].minimum_band_purity, descs[p].minimum_active_frame_ratio,
                                 descs[p].minimum_longest_active_run, descs[p].minimum_active_frame_mean_purity,
                                 descs[p].maximum_min_flank_purity, descs[p].maximum_max_flank_purity};
        for (int k = 0; k < 6; ++k) h_thr[p * 6 + k] = given[k] == given[k] ? given[k] : defaults[k];
        if (cl.strategy == APD_STRATEGY_MARKER_TONE && cl.tone_hz == cl.tone_hz) {
            int P = 2;
            // chirp-z over the K = L/2 + 1 bins the metrics use (not all L): the circular convolution only has to
            // hold indices k - n in (-L, K), so P >= L + K - 1 suffices (half the 2L - 1 of a full Bluestein DFT
            // for about half of all clip lengths)
            while (P < L + (L / 2 + 1) - 1) P <<= 1;
            cl.tone_P = P;
            max_P = std::max(max_P, P);
            std::vector<double2> tw(P / 2);
            for (int t = 0; t < P / 2; ++t) {
                const double a = -2.0 * M_PI * (double)t / (double)P;
                tw[t] = make_double2(cos(a), sin(a));
            }
            CK(upload(&cl.d_tone_tw, tw));
            CK(dalloc(&cl.d_tone_chirp, (size_t)P));
            CK(dalloc(&cl.d_tone_pre, (size_t)L));
            CK(dalloc(&cl.d_tone_post, (size_t)L / 2 + 1));
            double2 *b0 = nullptr, *b1 = nullptr;
            CK(dalloc(&b0, (size_t)P));
            CK(dalloc(&b1, (size_t)P));
            launch_tone_tables(L, P, cl.d_tone_tw, b0, b1, cl.d_tone_chirp, cl.d_tone_pre, cl.d_tone_post, 0);
            CK(cudaDeviceSynchronize());
            cudaFree(b0);
            cudaFree(b1);
            ++c->launches;
            h_P[p] = P; h_chirp[p] = cl.d_tone_chirp; h_tw[p] = cl.d_tone_tw;
            h_pre[p] = cl.d_tone_pre; h_post[p] = cl.d_tone_post;
        }
    }
    CK(cudaDeviceSynchronize());
    CK(upload(&c->d_clip_len, h_len));
    CK(upload(&c->d_clip_group, h_group));
    CK(upload(&c->d_strategy, h_strat));
    CK(upload(&c->d_is_short, h_short));
    CK(upload(&c->d_win_lo, h_lo));
    CK(upload(&c->d_win_hi, h_hi));
    CK(upload(&c->d_win_ds, h_ds));
    CK(upload(&c->d_tone_P, h_P));
    CK(upload(&c->d_self_max, h_smax));
    CK(upload(&c->d_self_corr_ptrs, h_cc));
    CK(upload(&c->d_win_cache_ptrs, h_cache));
    CK(upload(&c->d_clip_spec_ptrs, h_spec));
    CK(upload(&c->d_tone_hz, h_thz));
    CK(upload(&c->d_tone_thr, h_thr));
    CK(upload(&c->d_tone_chirp_ptrs, h_chirp));
    CK(upload(&c->d_tone_tw_ptrs, h_tw));
    CK(upload(&c->d_tone_pre_ptrs, h_pre));
    CK(upload(&c->d_tone_post_ptrs, h_post));
    cudaFree(tmp.max_bits); cudaFree(tmp.ints); cudaFree(tmp.zero_f); cudaFree((void*)tmp.spec_ptr);
    cudaFree(tmp.zero_ll); cudaFree(tmp.geom); cudaFree(tmp.desc);
    cudaFree(d_tl); cudaFree(d_tg); cudaFree(d_tstate); cudaFree(d_ten); cudaFree(d_tem1); cudaFree(d_tpatch);
    {
        std::vector<long long> h_off(n_clips);
        for (int p = 0; p < n_clips; ++p) h_off[p] = c->groups[c->clips[p].group].spec_off;
        CK(upload(&c->d_clip_spec_off, h_off));
    }

    // ---- workspace
    const int B = c->maxB;
    if (const char* e = getenv("APD_B200_SETS")) c->n_sets = std::min(3, std::max(2, atoi(e)));
    for (auto& b : used_sets(c)) {
        CK(dalloc(&b.d_geoms, (size_t)G));
        CK(cudaMallocHost((void**)&b.h_geoms, sizeof(SectionGeom) * G));
    }
    long long max_sec = c->C;
    for (auto& g : c->groups) max_sec = std::max(max_sec, c->C + g.halo);
    c->cells_stride = (int)((max_sec + c->kw.cell - 1) / c->kw.cell + 1);
    if (c->kw.general) c->cells_stride = (int)(G * 2 * ((max_sec + c->kw.cell - 1) / c->kw.cell + 4));   // 2 block slots per group
    CK(dalloc(&c->d_kw_state, (size_t)B * c->cells_stride * 4));
    CK(dalloc(&c->d_kw_energy, (size_t)B * c->cells_stride));
    CK(dalloc(&c->d_kw_em1, (size_t)B));
    CK(dalloc(&c->d_kw_patch, (size_t)B * G * c->kw.patch_cells));
    for (auto& b : used_sets(c)) {
        CK(dalloc(&b.d_lufs, (size_t)B * G));
        CK(dalloc(&b.d_gain, (size_t)B * G));
        CK(dalloc(&b.d_spec, (size_t)B * c->spec_slab));
    }
    c->n_slots = 128;                     // selected units per phase-2 round (further rounds if a batch selects more)
    if (const char* e = getenv("APD_B200_SLOTS")) c->n_slots = std::min(std::max(8, atoi(e)), (int)kMaxSlots);
    c->inv_units = 512;
    if (const char* e = getenv("APD_B200_INV_UNITS")) c->inv_units = std::max(1, atoi(e));
    c->scratch_elems = std::max<long long>((long long)B * (long long)max_class_groups, c->inv_units) * max_M;
    CK(dalloc(&c->d_scratch, (size_t)c->scratch_elems));
    c->d_scratch_fwd = c->d_scratch;
    if (const char* e = getenv("APD_B200_FWD_STREAM")) if (atoi(e)) {
        CK(dalloc(&c->d_scratch_fwd, (size_t)B * (size_t)max_class_groups * (size_t)max_M));
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&c->fwd, cudaStreamNonBlocking, atoi(e) == 2 ? lo : hi));
    }
    CK(cudaMalloc(&c->d_unit_desc, corr_inv_desc_bytes(c->inv_units)));
    CK(dalloc(&c->d_scratch_b, (size_t)c->inv_units * max_M));
    CK(cudaMalloc(&c->d_unit_desc_b, corr_inv_desc_bytes(c->inv_units)));
    CK(dalloc(&c->d_scratch2, (size_t)c->n_slots * max_M));
    CK(cudaMalloc(&c->d_unit_desc2, corr_inv_desc_bytes(c->n_slots)));
    c->sel_capacity = B * n_clips;
    for (auto& b : used_sets(c)) {
        CK(dalloc(&b.d_unit_max, (size_t)B * n_clips));
        CK(dalloc(&b.d_unit_npeaks, (size_t)B * n_clips));
        CK(dalloc(&b.d_counts, (size_t)S + 4));
        CK(cudaMallocHost((void**)&b.h_counts, sizeof(int) * (S + 4)));
        CK(dalloc(&b.d_sel, (size_t)c->sel_capacity));
        CK(cudaEventCreateWithFlags(&b.p1_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b.fwd_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b.loud_done, cudaEventDisableTiming));
    }
    {
        // both helper streams carry small latency-bound kernels: give them priority over the caller's stream so
        // their CTAs are placed as soon as a correlate CTA retires instead of queueing behind its whole grid
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        if (const char* e = getenv("APD_B200_PRIO")) if (!atoi(e)) hi = lo;
        CK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&c->pre, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&c->tone, cudaStreamNonBlocking, lo));
        CK(cudaStreamCreateWithPriority(&c->corr2, cudaStreamNonBlocking, lo));
        CK(cudaEventCreateWithFlags(&c->corr_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->corr_join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->tone_go, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->tone_done, cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&c->scan_start, cudaEventDisableTiming));
    c->cur_set = 0;
    use_set(c, 0);
    c->corr_stride = (max_nout + 31) / 32 * 32;
    c->cand_stride = max_nout / 2 + 2;
    if (max_nout / std::max(min_L, 1) + 2 > 8192)
        return fail(APD_ERR_UNSUPPORTED, "shortest clip too short for this chunk length (more than 8190 peaks per unit)");
    c->peak_stride = (int)(max_nout / std::max(min_L, 1) + 2);
    CK(dalloc(&c->d_corr, (size_t)c->n_slots * c->corr_stride));
    CK(dalloc(&c->d_cand_idx, (size_t)c->n_slots * c->cand_stride));
    CK(dalloc(&c->d_cand_val, (size_t)c->n_slots * c->cand_stride));
    CK(dalloc(&c->d_cand_state, (size_t)c->n_slots * c->cand_stride));
    CK(dalloc(&c->d_peaks, (size_t)c->n_slots * c->peak_stride));
    CK(dalloc(&c->d_peak_height, (size_t)c->n_slots * c->peak_stride));
    CK(dalloc(&c->d_n_peaks, (size_t)c->n_slots));
    CK(dalloc(&c->d_n_cands, (size_t)c->n_slots));
    CK(dalloc(&c->d_slot_cands, (size_t)c->n_slots * c->peak_stride));
    // records: an average of 4 per unit of a sub-batch, and never less than the worst case of ONE chunk (a unit keeps at
    // most N_out / L + 1 peaks, lib.rs:437-485), so that the caller can always make progress by scanning fewer chunks
    long long chunk_bound = 0, tone_bound = 0;
    for (auto& cl : c->clips) {
        const long long per_unit = max_nout / std::max(cl.L, 1) + 2;
        chunk_bound += per_unit;
        if (cl.tone_P > 0) tone_bound += per_unit;
    }
    if (chunk_bound > (1LL << 24)) return fail(APD_ERR_UNSUPPORTED, "clips too short for this chunk length (candidate bound)");
    c->out_capacity = (int)std::max<long long>(std::max(4096, B * n_clips * 4), chunk_bound);
    for (auto& b : used_sets(c)) {
        CK(dalloc(&b.d_out, (size_t)c->out_capacity));
        CK(cudaMallocHost((void**)&b.h_out, sizeof(apd_candidate) * (size_t)c->out_capacity));
        CK(cudaEventCreateWithFlags(&b.out_ready, cudaEventDisableTiming));
    }
    c->d_out = c->sets[c->cur_set].d_out;
    if (max_P > 0) {
        c->tone_stride = max_P;
        const long long per_cta = 3LL * 2 * max_P * (long long)sizeof(double2);
        c->tone_ctas = (int)std::max<long long>(1, std::min<long long>(64, (1LL << 31) / per_cta));
        if (const char* e = getenv("APD_B200_TONE_ITEMS")) c->tone_ctas = std::max(1, std::min(c->tone_ctas, atoi(e)));
        CK(dalloc(&c->d_tone_scratch, (size_t)c->tone_ctas * 3 * 2 * max_P));
        CK(dalloc(&c->d_tone_stats, (size_t)c->tone_ctas * 3 * 4));
        c->tone_max_P = max_P;
        for (auto& cl : c->clips)
            if (cl.tone_P > 0) c->tone_max_L = std::max(c->tone_max_L, cl.L);
        c->tone_item_cap = (int)std::max<long long>(std::max(4096, B * n_clips), tone_bound);
        for (auto& b : used_sets(c)) CK(cudaMalloc(&b.d_tone_items, tone_item_bytes() * c->tone_item_cap));
        c->d_tone_items = c->sets[c->cur_set].d_tone_items;
        CK(dalloc(&c->d_tone_metrics, (size_t)c->tone_item_cap * 15));
        CK(dalloc(&c->d_tone_alive, (size_t)c->tone_item_cap));
        const double wlr = std::nearbyint(0.025 * (double)sample_rate);
        c->tone_wl = wlr > 32.0 ? (int)wlr : 32;
    }
    CK(cudaDeviceSynchronize());
    return APD_OK;
}

extern "C" int apd_create(apd_ctx** out, int device, int sample_rate, int64_t chunk_samples, float height_min,
                          int n_clips, const apd_clip_desc* descs, int max_batch_chunks)
{
    if (!out || !descs || n_clips <= 0 || chunk_samples <= 0 || max_batch_chunks <= 0)
        return fail(APD_ERR_INVALID, "apd_create: bad arguments");
    apd_ctx* c = new apd_ctx();
    const int rc = create_ctx(c, device, sample_rate, chunk_samples, height_min, n_clips, descs, max_batch_chunks);
    if (rc != APD_OK) {
        const std::string msg = apd_last_error();          // apd_destroy must not lose the reason
        apd_destroy(c);                                     // frees whatever was allocated before the failure
        return fail(rc, msg);
    }
    *out = c;
    return APD_OK;
}

extern "C" int apd_destroy(apd_ctx* c)
{
    if (!c) return APD_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& cl : c->clips) {
        cudaFree(cl.d_raw); cudaFree(cl.d_norm); cudaFree(cl.d_rev); cudaFree(cl.d_spec);
        cudaFree(cl.d_self_corr); cudaFree(cl.d_win_cache); cudaFree(cl.d_tone_chirp); cudaFree(cl.d_tone_tw);
        cudaFree(cl.d_tone_pre); cudaFree(cl.d_tone_post);
    }
    for (auto& sc : c->shapes) {
        free_plan(&sc.plan);
        cudaFree(sc.d_clips); cudaFree(sc.d_group_halo); cudaFree(sc.d_group_index); cudaFree(sc.d_group_spec_off);
    }
    for (auto& kv : c->self_plans) free_plan(&kv.second);
    kw_config_destroy(&c->kw);
    for (auto& b : c->sets) {
        void* sp[] = {b.d_spec, b.d_unit_max, b.d_unit_npeaks, b.d_counts, b.d_sel, b.d_lufs, b.d_gain, b.d_geoms};
        for (void* q : sp) cudaFree(q);
        cudaFreeHost(b.h_counts);
        cudaFreeHost(b.h_geoms);
        cudaFree(b.d_out);
        cudaFree(b.d_tone_items);
        cudaFreeHost(b.h_out);
        if (b.out_ready) cudaEventDestroy(b.out_ready);
        if (b.p1_done) cudaEventDestroy(b.p1_done);
        if (b.fwd_done) cudaEventDestroy(b.fwd_done);
        if (b.loud_done) cudaEventDestroy(b.loud_done);
        for (auto& e : b.ev)
            if (e) cudaEventDestroy(e);
    }
    if (c->fwd) cudaStreamDestroy(c->fwd);
    if (c->d_scratch_fwd != c->d_scratch) cudaFree(c->d_scratch_fwd);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->pre) cudaStreamDestroy(c->pre);
    if (c->tone) cudaStreamDestroy(c->tone);
    if (c->corr2) cudaStreamDestroy(c->corr2);
    if (c->corr_fork) cudaEventDestroy(c->corr_fork);
    if (c->corr_join) cudaEventDestroy(c->corr_join);
    cudaFree(c->d_scratch_b);
    cudaFree(c->d_unit_desc_b);
    if (c->tone_go) cudaEventDestroy(c->tone_go);
    if (c->tone_done) cudaEventDestroy(c->tone_done);
    if (c->scan_start) cudaEventDestroy(c->scan_start);
    void* ptrs[] = {c->d_scratch2, c->d_unit_desc2,
                    c->d_clip_len, c->d_clip_group, c->d_strategy, c->d_is_short, c->d_win_lo, c->d_win_hi,
                    c->d_win_ds, c->d_tone_P, c->d_self_max, (void*)c->d_self_corr_ptrs, (void*)c->d_win_cache_ptrs,
                    (void*)c->d_clip_spec_ptrs, c->d_tone_hz, c->d_tone_thr, (void*)c->d_tone_chirp_ptrs,
                    (void*)c->d_tone_tw_ptrs, (void*)c->d_tone_pre_ptrs, (void*)c->d_tone_post_ptrs, c->d_peak_height,
                    c->d_clip_spec_off, c->d_kw_patch,
                    c->d_kw_state, c->d_kw_energy, c->d_kw_em1, c->d_scratch, c->d_corr,
                    c->d_cand_idx, c->d_cand_val, c->d_cand_state, c->d_peaks, c->d_n_peaks, c->d_n_cands,
                    c->d_slot_cands, c->d_tone_scratch, c->d_tone_metrics, c->d_tone_stats, c->d_tone_alive,
                    c->d_unit_desc};
    for (void* p : ptrs) cudaFree(p);
    delete c;
    return APD_OK;
}

extern "C" int apd_clip_info(apd_ctx* c, int clip, int32_t* sw, double* lufs, float* self_max, int32_t* fft_points)
{
    if (!c || clip < 0 || clip >= c->n_clips) return fail(APD_ERR_INVALID, "apd_clip_info: bad clip");
    const ClipHost& cl = c->clips[clip];
    if (sw) *sw = cl.sw;
    if (lufs) *lufs = cl.lufs;
    if (self_max) *self_max = cl.self_max;
    if (fft_points) *fft_points = 2 * c->shapes[c->groups[cl.group].shape].plan.M;
    return APD_OK;
}

extern "C" int apd_clip_normalized(apd_ctx* c, int clip, float* out_host)
{
    if (!c || clip < 0 || clip >= c->n_clips || !out_host) return fail(APD_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(out_host, c->clips[clip].d_norm, sizeof(float) * c->clips[clip].L, cudaMemcpyDeviceToHost));
    return APD_OK;
}

extern "C" int apd_clip_self_correlation(apd_ctx* c, int clip, float* out_host)
{
    if (!c || clip < 0 || clip >= c->n_clips || !out_host) return fail(APD_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(out_host, c->clips[clip].d_self_corr, sizeof(float) * (2 * c->clips[clip].L - 1),
                  cudaMemcpyDeviceToHost));
    return APD_OK;
}

// ---------------------------------------------------------------------------
// stages
// ---------------------------------------------------------------------------
static int stage_begin(apd_ctx* c, const float* audio, int64_t base, int64_t n, int cb, int ce, cudaStream_t st)
{
    if (!c || !audio || n <= 0 || ce <= cb || cb < 0) return fail(APD_ERR_INVALID, "scan: bad arguments");
    if (ce - cb > c->maxB) return fail(APD_ERR_INVALID, "scan: more chunks than max_batch_chunks");
    if (cb > 0 && (long long)cb * c->C - c->max_halo < base)
        return fail(APD_ERR_INVALID, "scan: audio region does not contain the look-back halo of chunk_begin");
    if (cb == 0 && base != 0) return fail(APD_ERR_INVALID, "scan: chunk 0 requires base_sample == 0");
    if ((long long)(ce - 1) * c->C >= base + n) return fail(APD_ERR_INVALID, "scan: chunk_end beyond the audio region");
    CK(cudaSetDevice(c->device));
    c->audio = audio; c->base = base; c->nsamp = n; c->chunk_begin = cb; c->chunk_end = ce;
    fill_geoms(c);
    CK(cudaMemcpyAsync(c->d_geoms, c->h_geoms, sizeof(SectionGeom) * c->groups.size(), cudaMemcpyHostToDevice, st));
    const int B = ce - cb;
    CK(cudaMemsetAsync(c->d_unit_max, 0, sizeof(unsigned int) * (size_t)B * c->n_clips, st));
    CK(cudaMemsetAsync(c->d_unit_npeaks, 0xff, sizeof(int) * (size_t)B * c->n_clips, st));
    CK(cudaMemsetAsync(c->d_counts, 0, sizeof(int) * (c->shapes.size() + 4), st));
    c->staged = true;
    return APD_OK;
}

static int stage_loudness(apd_ctx* c, cudaStream_t st)
{
    const int B = c->chunk_end - c->chunk_begin, G = (int)c->groups.size();
    launch_loudness(c->kw, union_geom(c), c->h_geoms, c->d_geoms, G, B, c->cells_stride, c->d_kw_state,
                    c->d_kw_energy, c->d_kw_em1, c->d_kw_patch, c->d_lufs, c->d_gain, st);
    c->launches += 5;
    CK(cudaGetLastError());
    return APD_OK;
}

static int stage_forward(apd_ctx* c, cudaStream_t st)
{
    const int B = c->chunk_end - c->chunk_begin, G = (int)c->groups.size();
    for (auto& sc : c->shapes) {
        if (corr_inv_supported(sc.plan)) {
            // all sliding-window groups of the shape in one pair of launches
            FwdGroups FG{sc.d_group_halo, sc.d_group_index, sc.d_group_spec_off, (int)sc.groups.size()};
            launch_forward(sc.plan, c->h_geoms[sc.groups[0]], FG, c->d_gain, G, B, c->d_scratch_fwd, c->d_spec,
                           c->spec_slab, st);
            c->launches += 2;
            continue;
        }
        for (int g : sc.groups) {
            launch_forward(sc.plan, c->h_geoms[g], FwdGroups{nullptr, nullptr, nullptr, 0}, c->d_gain + g, G, B,
                           c->d_scratch_fwd, c->d_spec + c->groups[g].spec_off, c->spec_slab, st);
            c->launches += 2;
        }
    }
    CK(cudaGetLastError());
    return APD_OK;
}

static int stage_correlate_max(apd_ctx* c, cudaStream_t st)
{
    const int B = c->chunk_end - c->chunk_begin;
    InvOut O{c->d_unit_max, c->n_clips, nullptr, 0, c->d_self_max};
    const UnitCtx X = unit_ctx(c);
    static const bool clip_major = !(getenv("APD_B200_CHUNK_MAJOR") && atoi(getenv("APD_B200_CHUNK_MAJOR")));
    // launches alternate between the caller's stream and a second one of the same priority (own intermediate
    // and descriptors), so that consecutive launches overlap
    static const bool two_streams = !(getenv("APD_B200_CORR_STREAMS") && atoi(getenv("APD_B200_CORR_STREAMS")) == 1);
    if (two_streams) {
        CK(cudaEventRecord(c->corr_fork, st));
        CK(cudaStreamWaitEvent(c->corr2, c->corr_fork, 0));
    }
    int launch = 0;
    for (auto& sc : c->shapes) {
        const int ns = (int)sc.clips.size();
        const long long nunits = (long long)B * ns;
        // launches of equal size (a multiple of 16 units) instead of full ones plus a small remainder
        const long long nl = (nunits + c->inv_units - 1) / c->inv_units;
        const long long step = std::min<long long>(c->inv_units, ((nunits + nl - 1) / nl + 15) / 16 * 16);
        for (long long u0 = 0; u0 < nunits; u0 += step, ++launch) {
            UnitSrc U{nullptr, nullptr, sc.d_clips, ns, (int)u0, clip_major ? B : 0, 0, 0};
            const bool odd = two_streams && (launch & 1);
            launch_inverse(sc.plan, X, c->d_spec, c->spec_slab, U, (int)std::min<long long>(step, nunits - u0),
                           odd ? c->d_scratch_b : c->d_scratch, odd ? c->d_unit_desc_b : c->d_unit_desc, O, false,
                           odd ? c->corr2 : st);
            c->launches += 3;
        }
    }
    if (two_streams) {
        CK(cudaEventRecord(c->corr_join, c->corr2));
        CK(cudaStreamWaitEvent(st, c->corr_join, 0));
    }
    CK(cudaGetLastError());
    return APD_OK;
}

// One phase-2 round: slots [slot0, slot0 + n_slots) of the selected list.
static void phase2_round(apd_ctx* c, int slot0, cudaStream_t st)
{
    const int S = (int)c->shapes.size(), G = (int)c->groups.size();
    const int ns = c->n_slots;
    const UnitCtx X = unit_ctx(c);
    InvOut O{c->d_unit_max, c->n_clips, c->d_corr, c->corr_stride, c->d_self_max};
    // timing experiments: 0 = selection only, 1 = + write-back inverse, 2 = + find_peaks, 3+ = everything
    static const int level = getenv("APD_B200_P2_LEVEL") ? atoi(getenv("APD_B200_P2_LEVEL")) : 9;
    if (level < 1) return;
    for (int s = 0; s < S; ++s) {
        UnitSrc U{c->d_sel, c->d_counts + s, nullptr, 0, slot0, 0, 0, 0};
        launch_inverse(c->shapes[s].plan, X, c->d_spec, c->spec_slab, U, ns, c->d_scratch2, c->d_unit_desc2, O, true, st);
        c->launches += 3;
    }
    if (level < 2) return;
    PeakArgs PA{c->d_sel, c->d_counts + S, slot0, c->d_geoms, c->d_clip_group, c->d_clip_len, c->height,
                c->d_corr, c->corr_stride, c->d_cand_idx, c->d_cand_val, c->d_cand_state, c->cand_stride,
                c->d_peaks, c->d_peak_height, c->peak_stride, c->d_n_peaks, c->d_n_cands, c->d_counts + S + 2};
    launch_find_peaks(PA, ns, st);
    ++c->launches;
    if (level < 3) return;
    ClipVerify CV{c->d_clip_len, c->d_clip_group, c->d_strategy, c->d_self_corr_ptrs, c->d_win_lo,
                  c->d_win_hi, c->d_win_ds, c->d_win_cache_ptrs, c->d_is_short, c->d_tone_hz, c->d_tone_thr,
                  c->d_tone_P, c->d_tone_chirp_ptrs, c->d_tone_tw_ptrs, c->d_tone_pre_ptrs, c->d_tone_post_ptrs};
    VerifyArgs VA{PA, CV, c->chunk_begin, c->sr, c->d_gain, G, c->d_out, c->d_counts + S + 1, c->out_capacity,
                  c->d_slot_cands, c->d_tone_scratch, c->tone_stride, c->tone_ctas};
    launch_verify(VA, ns, st, &c->launches);
    if (c->tone_ctas > 0)
        launch_tone_collect(VA, ns, c->d_tone_items, c->d_counts + S + 3, c->tone_item_cap, st, &c->launches);
    launch_emit(VA, ns, st, &c->launches);
    k_record_npeaks<<<(ns + 127) / 128, 128, 0, st>>>(c->d_sel, c->d_counts + S, slot0, ns, c->d_n_peaks,
                                                     c->n_clips, c->d_unit_npeaks);
    ++c->launches;
}

static void phase2_tone(apd_ctx* c, int n_items, cudaStream_t st)
{
    if (c->tone_ctas <= 0 || n_items <= 0) return;
    const int S = (int)c->shapes.size(), G = (int)c->groups.size();
    PeakArgs PA{c->d_sel, c->d_counts + S, 0, c->d_geoms, c->d_clip_group, c->d_clip_len, c->height,
                c->d_corr, c->corr_stride, c->d_cand_idx, c->d_cand_val, c->d_cand_state, c->cand_stride,
                c->d_peaks, c->d_peak_height, c->peak_stride, c->d_n_peaks, c->d_n_cands, c->d_counts + S + 2};
    ClipVerify CV{c->d_clip_len, c->d_clip_group, c->d_strategy, c->d_self_corr_ptrs, c->d_win_lo,
                  c->d_win_hi, c->d_win_ds, c->d_win_cache_ptrs, c->d_is_short, c->d_tone_hz, c->d_tone_thr,
                  c->d_tone_P, c->d_tone_chirp_ptrs, c->d_tone_tw_ptrs, c->d_tone_pre_ptrs, c->d_tone_post_ptrs};
    VerifyArgs VA{PA, CV, c->chunk_begin, c->sr, c->d_gain, G, c->d_out, c->d_counts + S + 1, c->out_capacity,
                  c->d_slot_cands, c->d_tone_scratch, c->tone_stride, c->tone_ctas};
    launch_tone_batch(VA, c->d_tone_items, c->d_counts + S + 3, n_items, c->d_tone_metrics, c->d_tone_stats,
                      c->tone_ctas, c->tone_max_P, c->tone_max_L, c->tone_wl, c->tone_all_segments, c->d_tone_alive, st,
                      &c->launches);
}

// Select the units that can have peaks, then run the first phase-2 round without waiting for the
// count (CTAs beyond it exit at once); collect() runs further rounds in the rare case that more than
// n_slots units of a batch were selected.
static int stage_peaks_verify(apd_ctx* c, cudaStream_t st)
{
    const int B = c->chunk_end - c->chunk_begin, S = (int)c->shapes.size();
    for (int s = 0; s < S; ++s) {
        ShapeClass& sc = c->shapes[s];
        k_select_shape<<<1, 1024, 0, st>>>(c->d_unit_max, c->d_self_max, sc.d_clips, (int)sc.clips.size(),
                                           c->n_clips, B, c->height, c->d_sel, c->d_counts + s, c->d_counts + s + 1,
                                           c->sel_capacity, c->d_counts + S + 2);
        ++c->launches;
    }
    phase2_round(c, 0, st);
    CK(cudaGetLastError());
    return APD_OK;
}

extern "C" int apd_stage_loudness(apd_ctx* c, const float* audio, int64_t base, int64_t n, int32_t cb, int32_t ce,
                                  void* stream)
{
    int rc = stage_begin(c, audio, base, n, cb, ce, (cudaStream_t)stream);
    if (rc) return rc;
    return stage_loudness(c, (cudaStream_t)stream);
}
extern "C" int apd_stage_forward_fft(apd_ctx* c, void* stream)
{
    if (!c || !c->staged) return fail(APD_ERR_INVALID, "no staged batch");
    return stage_forward(c, (cudaStream_t)stream);
}
extern "C" int apd_stage_correlate_max(apd_ctx* c, void* stream)
{
    if (!c || !c->staged) return fail(APD_ERR_INVALID, "no staged batch");
    return stage_correlate_max(c, (cudaStream_t)stream);
}
extern "C" int apd_stage_peaks_verify(apd_ctx* c, void* stream)
{
    if (!c || !c->staged) return fail(APD_ERR_INVALID, "no staged batch");
    return stage_peaks_verify(c, (cudaStream_t)stream);
}

static bool cand_less(const apd_candidate& a, const apd_candidate& b)
{
    if (a.chunk != b.chunk) return a.chunk < b.chunk;
    if (a.clip != b.clip) return a.clip < b.clip;
    return a.peak < b.peak;
}

// Finish phase 2 of the batch in the current set, in two halves so that the marker-tone verification of one
// sub-batch overlaps the normal verification of the next:
//   collect_begin (side stream st): wait for the first phase-2 round, run further rounds if more than n_slots units
//     were selected, then enqueue the deferred tone verification and the copy of records + counters to pinned
//     host memory on the tone stream (every tone work item appends exactly one record, so the record count is
//     known before the tone kernels run); records the set's out_ready event.
//   collect_end: wait for out_ready, append the records to cand_host[*n_cand ...] and fill the optional traces.
static int collect_begin(apd_ctx* c, cudaStream_t st)
{
    const int S = (int)c->shapes.size();
    apd_ctx::BatchSet& set = c->sets[c->cur_set];
    CK(cudaMemcpyAsync(c->h_counts, c->d_counts, sizeof(int) * (S + 4), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (c->h_counts[S] > c->n_slots) {
        for (int slot0 = c->n_slots; slot0 < c->h_counts[S]; slot0 += c->n_slots) phase2_round(c, slot0, st);
        CK(cudaMemcpyAsync(c->h_counts, c->d_counts, sizeof(int) * (S + 4), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    // the side stream is idle here (the host has just synchronised it), so the tone stream needs no event to
    // see the work items; APD_B200_TONE_LOW=0 keeps the tone kernels on the side stream itself
    static const bool skip_tone = getenv("APD_B200_SKIP_TONE") && atoi(getenv("APD_B200_SKIP_TONE"));   // timing experiments
    static const bool tone_low = !(getenv("APD_B200_TONE_LOW") && !atoi(getenv("APD_B200_TONE_LOW")));
    cudaStream_t ts = (tone_low && st == c->side) ? c->tone : st;
    const int n_tone = skip_tone ? 0 : c->h_counts[S + 3];
    c->work[0] += c->h_counts[S];
    c->work[1] += c->h_counts[S + 1] + n_tone;
    c->work[2] += n_tone;
    c->work[3] += 1;
    if (n_tone > 0) phase2_tone(c, n_tone, ts);
    if (c->profile) cudaEventRecord(c->ev[6], ts);
    set.n_expected = std::min(c->h_counts[S + 1] + n_tone, c->out_capacity);
    CK(cudaMemcpyAsync(c->h_counts, c->d_counts, sizeof(int) * (S + 4), cudaMemcpyDeviceToHost, ts));
    if (set.n_expected > 0)
        CK(cudaMemcpyAsync(set.h_out, c->d_out, sizeof(apd_candidate) * (size_t)set.n_expected, cudaMemcpyDeviceToHost, ts));
    CK(cudaEventRecord(set.out_ready, ts));
    return APD_OK;
}

static int collect_end(apd_ctx* c, apd_candidate* cand_host, int32_t cap, int32_t* n_cand, apd_unit_trace* trace,
                       double* lufs_host)
{
    const int B = c->chunk_end - c->chunk_begin, G = (int)c->groups.size(), S = (int)c->shapes.size();
    apd_ctx::BatchSet& set = c->sets[c->cur_set];
    cudaStream_t st = c->side;
    CK(cudaEventSynchronize(set.out_ready));
    const int n = c->h_counts[S + 1];
    if (c->h_counts[S + 2]) return fail(APD_ERR_OVERFLOW, "candidate workspace overflow (flags " +
                                        std::to_string(c->h_counts[S + 2]) + ")");
    if (n > set.n_expected) return fail(APD_ERR_OVERFLOW, "more candidate records than work items");
    if (*n_cand + n > cap) return fail(APD_ERR_OVERFLOW, "candidate buffer too small");
    apd_candidate* dst = cand_host + *n_cand;
    if (n > 0) memcpy(dst, set.h_out, sizeof(apd_candidate) * (size_t)n);
    std::vector<unsigned int> um;
    std::vector<int> np;
    if (trace) {
        um.resize((size_t)B * c->n_clips);
        np.resize((size_t)B * c->n_clips);
        CK(cudaMemcpyAsync(um.data(), c->d_unit_max, sizeof(unsigned int) * um.size(), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(np.data(), c->d_unit_npeaks, sizeof(int) * np.size(), cudaMemcpyDeviceToHost, st));
    }
    std::vector<double> lf;
    if (lufs_host) {
        lf.resize((size_t)B * G);
        CK(cudaMemcpyAsync(lf.data(), c->d_lufs, sizeof(double) * lf.size(), cudaMemcpyDeviceToHost, st));
    }
    if (trace || lufs_host) CK(cudaStreamSynchronize(st));      // (waits for whatever phase 2 the side stream is running)
    std::sort(dst, dst + n, cand_less);
    *n_cand += n;
    if (trace) {
        for (int ci = 0; ci < B; ++ci)
            for (int p = 0; p < c->n_clips; ++p) {
                apd_unit_trace& t = trace[(size_t)ci * c->n_clips + p];
                float am;
                memcpy(&am, &um[(size_t)ci * c->n_clips + p], sizeof(float));
                t.absmax = am;
                t.max_choose = std::max(c->clips[p].self_max, am);
                long long s;
                int ns;
                section_bounds(c->h_geoms[c->clips[p].group], ci, s, ns);
                t.n_out = ns > 0 ? ns + c->clips[p].L - 1 : 0;
                t.n_peaks = np[(size_t)ci * c->n_clips + p];
            }
    }
    if (lufs_host)
        for (int ci = 0; ci < B; ++ci)
            for (int p = 0; p < c->n_clips; ++p)
                lufs_host[(size_t)ci * c->n_clips + p] = lf[(size_t)ci * G + c->clips[p].group];
    if (c->profile) {
        const int pairs[4][2] = {{0, 1}, {2, 3}, {7, 4}, {5, 6}};
        for (int i = 0; i < 4; ++i) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, c->ev[pairs[i][0]], c->ev[pairs[i][1]]) == cudaSuccess) c->stage_ms[i] += ms;
        }
    }
    return APD_OK;
}

// Scan = for each sub-batch of at most max_batch_chunks chunks:
//   phase 1 on the caller's stream : loudness -> forward FFT -> fused correlate + max          (all units)
//   phase 2 on the side stream     : selection -> write-back inverse -> peaks -> verification  (selected units)
// Phase 2 of sub-batch k overlaps phase 1 of sub-batch k + 1; what phase 1 hands over is double-buffered.
extern "C" int apd_scan(apd_ctx* c, const float* audio, int64_t base, int64_t n, int32_t cb, int32_t ce,
                        apd_candidate* cand_host, int32_t cap, int32_t* n_cand, apd_unit_trace* trace,
                        double* lufs_host, void* stream)
{
    cudaStream_t s1 = (cudaStream_t)stream;
    if (!c || !cand_host || !n_cand) return fail(APD_ERR_INVALID, "scan: null argument");
    if (ce <= cb) return fail(APD_ERR_INVALID, "scan: empty chunk range");
    CK(cudaSetDevice(c->device));
    // with a trace every tone metric is reported; without, the flanks of candidates whose matched segment already
    // fails are not computed (same decisions, verify.cu: k_tone_gate)
    c->tone_all_segments = trace != nullptr || (getenv("APD_B200_TONE_ALL") && atoi(getenv("APD_B200_TONE_ALL")));
    static const bool loud_stream = !(getenv("APD_B200_LOUD_STREAM") && !atoi(getenv("APD_B200_LOUD_STREAM")));
    cudaStream_t s2 = c->side, s3 = loud_stream ? c->pre : s1;
    *n_cand = 0;
    // the loudness stream must see what the caller enqueued before the scan (e.g. the upload of the audio)
    CK(cudaEventRecord(c->scan_start, s1));
    CK(cudaStreamWaitEvent(s3, c->scan_start, 0));
    const int nb = (ce - cb + c->maxB - 1) / c->maxB;
    const bool prof = c->profile;
    int rc = APD_OK;
    auto phase1 = [&](int k) -> int {
        use_set(c, k % c->n_sets);
        const int b0 = cb + k * c->maxB, b1 = std::min<int>(ce, b0 + c->maxB);
        // loudness on its own stream: enqueued while the previous sub-batch's correlate stage is still running
        int r = stage_begin(c, audio, base, n, b0, b1, s3);
        if (r) return r;
        if (prof) cudaEventRecord(c->ev[0], s3);
        if ((r = stage_loudness(c, s3))) return r;
        if (prof) cudaEventRecord(c->ev[1], s3);
        CK(cudaEventRecord(c->sets[k % c->n_sets].loud_done, s3));
        cudaStream_t sf = c->fwd ? c->fwd : s1;
        CK(cudaStreamWaitEvent(sf, c->sets[k % c->n_sets].loud_done, 0));
        if (prof) cudaEventRecord(c->ev[2], sf);
        if ((r = stage_forward(c, sf))) return r;
        if (prof) cudaEventRecord(c->ev[3], sf);
        if (c->fwd) {
            CK(cudaEventRecord(c->sets[k % c->n_sets].fwd_done, sf));
            CK(cudaStreamWaitEvent(s1, c->sets[k % c->n_sets].fwd_done, 0));
        }
        if (prof) cudaEventRecord(c->ev[7], s1);
        if ((r = stage_correlate_max(c, s1))) return r;
        if (prof) cudaEventRecord(c->ev[4], s1);
        CK(cudaEventRecord(c->sets[k % c->n_sets].p1_done, s1));
        return APD_OK;
    };
    auto phase2_begin = [&](int k) -> int {
        use_set(c, k % c->n_sets);
        CK(cudaStreamWaitEvent(s2, c->sets[k % c->n_sets].p1_done, 0));
        if (prof) cudaEventRecord(c->ev[5], s2);
        static const bool skip_p2 = getenv("APD_B200_SKIP_P2") && atoi(getenv("APD_B200_SKIP_P2"));   // timing experiments
        if (skip_p2) return APD_OK;
        return stage_peaks_verify(c, s2);
    };
    auto phase2_collect = [&](int k) -> int {
        use_set(c, k % c->n_sets);
        return collect_begin(c, s2);
    };
    auto phase2_end = [&](int k) -> int {
        use_set(c, k % c->n_sets);
        const size_t off = (size_t)(c->chunk_begin - cb) * c->n_clips;
        return collect_end(c, cand_host, cap, n_cand, trace ? trace + off : nullptr, lufs_host ? lufs_host + off : nullptr);
    };
    // Host order.  With two sets (default) phase 2 of sub-batch k - 1, tone included, overlaps phase 1 of k and is
    // finished before phase 2 of k starts.  With three sets (APD_B200_SETS=3) the marker-tone verification of k - 1
    // (tone stream) also overlaps the normal phase 2 of k (side stream) and phase 1 of k + 1; the host collects the
    // records of k - 2.  Measured equal: the GPU is work-conserving, phase 2 costs its SM / HBM time either way.
    auto bail = [&](int r) { cudaDeviceSynchronize(); return r; };
    if (c->n_sets == 3) {
        for (int k = 0; k < std::min(2, nb); ++k)
            if ((rc = phase1(k))) return bail(rc);
        if ((rc = phase2_begin(0))) return bail(rc);
        for (int k = 1; k < nb; ++k) {
            if ((rc = phase2_collect(k - 1)) || (rc = phase2_begin(k))) return bail(rc);
            if (k >= 2 && (rc = phase2_end(k - 2))) return bail(rc);
            if (k + 1 < nb && (rc = phase1(k + 1))) return bail(rc);          // its set was that of k - 2
        }
        if ((rc = phase2_collect(nb - 1))) return bail(rc);
        if (nb >= 2 && (rc = phase2_end(nb - 2))) return bail(rc);
        rc = phase2_end(nb - 1);
    } else {
        if ((rc = phase1(0)) || (rc = phase2_begin(0))) return bail(rc);
        for (int k = 1; k < nb; ++k) {
            // enqueue the next sub-batch's phase 1 first so the GPU stays busy while the host waits for phase 2
            if ((rc = phase1(k)) || (rc = phase2_collect(k - 1)) || (rc = phase2_end(k - 1)) || (rc = phase2_begin(k)))
                return bail(rc);
        }
        if (!(rc = phase2_collect(nb - 1))) rc = phase2_end(nb - 1);
    }
    if (rc) cudaDeviceSynchronize();
    return rc;
}

extern "C" int apd_stage_unit_correlation(apd_ctx* c, int32_t chunk, int32_t clip, float* out_host, int32_t capacity,
                                          int32_t* n_out, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!c || !c->staged || clip < 0 || clip >= c->n_clips || chunk < c->chunk_begin || chunk >= c->chunk_end)
        return fail(APD_ERR_INVALID, "unit_correlation: bad unit");
    const ClipHost& cl = c->clips[clip];
    const Fft4Plan& plan = c->shapes[c->groups[cl.group].shape].plan;
    long long s;
    int ns;
    section_bounds(c->h_geoms[cl.group], chunk - c->chunk_begin, s, ns);
    const int no = ns > 0 ? ns + cl.L - 1 : 0;
    if (no > capacity) return fail(APD_ERR_INVALID, "unit_correlation: buffer too small");
    int2 u = make_int2(chunk - c->chunk_begin, clip);
    int2* d_u = nullptr;
    int* d_rng = nullptr;
    const int rng[2] = {0, 1};
    CK(dalloc(&d_u, 1));
    CK(dalloc(&d_rng, 2));
    CK(cudaMemcpyAsync(d_u, &u, sizeof(u), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_rng, rng, sizeof(rng), cudaMemcpyHostToDevice, st));
    UnitSrc U{d_u, d_rng, nullptr, 0, 0, 0, 0, 0};
    InvOut O{c->d_unit_max, c->n_clips, c->d_corr, c->corr_stride, c->d_self_max};
    launch_inverse(plan, unit_ctx(c), c->d_spec, c->spec_slab, U, 1, c->d_scratch2, c->d_unit_desc2, O, true, st);
    c->launches += 3;
    CK(cudaMemcpyAsync(out_host, c->d_corr, sizeof(float) * no, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_u);
    cudaFree(d_rng);
    *n_out = no;
    return APD_OK;
}

// Marker-tone verification of ONE candidate of an arbitrary, already normalised section (reference
// audio_pattern_detector.py:660-750, _verify_marker_tone(audio_section, peak, ...)): the section is treated as chunk 0
// of a stream of its own with unit gain, the candidate becomes the only tone work item, and the device kernels of a
// scan compute the 3 x 5 metrics and the accept decision.  For the reference's verifier tests and known answers.
extern "C" int apd_verify_tone(apd_ctx* c, int32_t clip, const float* section_dev, int32_t n, int32_t peak,
                               double* metrics15_host, int32_t* accept, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!c || clip < 0 || clip >= c->n_clips || !section_dev || n <= 0 || !metrics15_host || !accept)
        return fail(APD_ERR_INVALID, "verify_tone: bad arguments");
    if (c->tone_ctas <= 0 || c->clips[clip].strategy != APD_STRATEGY_MARKER_TONE || c->clips[clip].tone_hz != c->clips[clip].tone_hz)
        return fail(APD_ERR_INVALID, "verify_tone: not a marker-tone clip");
    if (n > c->C) return fail(APD_ERR_INVALID, "verify_tone: section longer than a chunk");
    use_set(c, 0);
    int rc = stage_begin(c, section_dev, 0, n, 0, 1, st);
    if (rc) return rc;
    const int G = (int)c->groups.size(), S = (int)c->shapes.size();
    std::vector<double> ones((size_t)G, 1.0);
    CK(cudaMemcpyAsync(c->d_gain, ones.data(), sizeof(double) * G, cudaMemcpyHostToDevice, st));
    struct { int ci, clip, peak; float height; } item = {0, clip, peak, 1.0f};
    static_assert(sizeof(item) == 16, "ToneItem layout");
    if (tone_item_bytes() != sizeof(item)) return fail(APD_ERR_INVALID, "verify_tone: work item layout");
    CK(cudaMemcpyAsync(c->d_tone_items, &item, sizeof(item), cudaMemcpyHostToDevice, st));
    const int one = 1;
    CK(cudaMemcpyAsync(c->d_counts + S + 3, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    c->tone_all_segments = true;
    phase2_tone(c, 1, st);
    apd_candidate rec;
    CK(cudaMemcpyAsync(&rec, c->d_out, sizeof(rec), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    c->staged = false;
    for (int sgm = 0; sgm < 3; ++sgm)
        for (int j = 0; j < 5; ++j) metrics15_host[sgm * 5 + j] = rec.tone[sgm][j];
    *accept = (rec.flags & APD_FLAG_ACCEPT) ? 1 : 0;
    return APD_OK;
}

extern "C" int64_t apd_launch_count(apd_ctx* c) { return c ? c->launches : 0; }

extern "C" int apd_profile(apd_ctx* c, int enable)
{
    if (!c) return fail(APD_ERR_INVALID, "apd_profile: null context");
    CK(cudaSetDevice(c->device));
    if (enable && !c->sets[0].ev[0])
        for (auto& b : used_sets(c))
            for (auto& e : b.ev) CK(cudaEventCreate(&e));
    c->profile = enable != 0;
    return APD_OK;
}

extern "C" int apd_profile_read(apd_ctx* c, double* ms4, int reset)
{
    if (!c || !ms4) return fail(APD_ERR_INVALID, "apd_profile_read: bad arguments");
    for (int i = 0; i < 4; ++i) {
        ms4[i] = c->stage_ms[i];
        if (reset) c->stage_ms[i] = 0.0;
    }
    return APD_OK;
}

extern "C" int apd_work_counters(apd_ctx* c, int64_t* out4, int reset)
{
    if (!c || !out4) return fail(APD_ERR_INVALID, "apd_work_counters: bad arguments");
    for (int i = 0; i < 4; ++i) {
        out4[i] = c->work[i];
        if (reset) c->work[i] = 0;
    }
    return APD_OK;
}

extern "C" int apd_unit_n_out(apd_ctx* c, int32_t chunk, int32_t clip, int64_t total_samples, int32_t* n_out)
{
    if (!c || clip < 0 || clip >= c->n_clips || !n_out) return fail(APD_ERR_INVALID, "bad arguments");
    SectionGeom G{nullptr, 0, total_samples, c->C, 0, c->groups[c->clips[clip].group].halo};
    long long s;
    int ns;
    section_bounds(G, chunk, s, ns);
    *n_out = ns > 0 ? ns + c->clips[clip].L - 1 : 0;
    return APD_OK;
}
