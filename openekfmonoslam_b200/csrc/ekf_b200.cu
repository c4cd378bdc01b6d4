// ekf_b200.cu -- context management and the C ABI of include/ekf_b200.h.
// Host side of the B200-native EKF hot path: owns device memory, the stream and the per-frame
// launch sequence (EKF::step order, E/EKF.cpp:242-572).  Two small device->host reads per frame
// (after RANSAC and after the rescue gate) size the update launches; everything else is
// asynchronous on the handle's stream.
#include "../../include/ekf_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ekf_linalg.cuh"
#include "ekf_schain.cuh"
#include "ekf_chain.cuh"
#include "ekf_map.cuh"
#include "ekf_ncc.cuh"
#include "ekf_frontend.cuh"
#include "ekf_downdate_tma.cuh"
#include "ekf_update_small.cuh"
#include "ekf_gemm_tma.cuh"

using namespace ekf;

static thread_local std::string g_err;

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char buf_[512];                                                                            \
            snprintf(buf_, sizeof(buf_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            g_err = buf_;                                                                              \
            return EKFB_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)

#define REQUIRE(cond, msg)          \
    do {                            \
        if (!(cond)) {              \
            g_err = msg;            \
            return EKFB_ERR_ARG;    \
        }                           \
    } while (0)

constexpr int kFusedSmemMax = 227 * 1024 - 1024;   // k_update_fused also has a few bytes of static shared memory

enum ProfGroup { G_PREDICT = 0, G_MEASURE, G_MATCH, G_RANSAC, G_GAIN, G_CHOL, G_DOWNDATE, G_RESCUE, G_MISC, G_COUNT };

struct SeqDev {
    float* xy = nullptr;
    uint8_t* desc = nullptr;
    std::vector<int> off;
};

// A lane = a contiguous range of the handle's filters that runs its frame on its own stream (batched handles only): the
// latency-bound phases of one lane (factorisation steps, RANSAC, the host's two reads of the counters) overlap the
// throughput-bound phases (covariance downdate) of the other.  See step_lanes.
struct Lane {
    int idx = 0, f0 = 0, F = 0;
    cudaStream_t stream = nullptr;   // lane 0 runs on the handle's stream
    cudaEvent_t done = nullptr;
    std::vector<int> hn, hN, hKp;
    DevView v;
    int seq = 0, chunk0 = 0;
    bool pending = false;
};

struct ekfb_ctx {
    ekfb_params prm;
    int device = 0, F = 0, Nmax = 0, nmax = 0, ld = 0, Kpmax = 0, kmax = 0, ldS = 0, supWords = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    int smCount = 148;
    DevView v;
    std::vector<void*> allocs;
    std::vector<int> hn, hN, hKp;  // host copies of per-filter sizes
    int* h_dims = nullptr;          // pinned mirror of v.dims
    int* h_dims_zc = nullptr;       // mapped pinned memory the counter-publishing kernels write (publish_dims)
    volatile int* h_flag_zc = nullptr;
    int dims_seq = 0;
    float* d_kpxy = nullptr;        // staging for ekfb_set_keypoints
    uint8_t* d_kpdesc = nullptr;
    const float** h_kpxy_ptr = nullptr;  // pinned
    const uint8_t** h_kpdesc_ptr = nullptr;
    const float** d_kpxy_ptr = nullptr;
    const uint8_t** d_kpdesc_ptr = nullptr;
    int* h_kpcount = nullptr;       // pinned
    int* d_kpcount = nullptr;
    std::vector<SeqDev> seq;
    RecordDev* d_rec = nullptr;
    RecordDev* h_rec = nullptr;  // pinned
    cudaEvent_t timers[64];
    // profiling
    bool prof = false;
    cudaEvent_t pe[2];
    float prof_ms[G_COUNT];
    int prof_launch[G_COUNT];
    int cur_group = G_MISC;
    int64_t launches = 0;
    float last_downdate_ms = 0.f;
    // L2 flush scratch and the per-launch event pool of the downdate kernel
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    int slab_trsm_max_k = 1152;   // single filter: above this many rows the TRSM on the global-memory resident B takes over from the
                                  // 16-column shared-memory slabs (option 17; measured equal at k = 1028: 3.005 vs 3.000 ms per C5 frame)
    int force_generic = 0;     // option 1: 1 = right-looking factorisation over the augmented matrix, 2 = S-chain + global-memory TRSM (ekf_gemm_tma.cuh)
    int downdate_variant = 2;   // 2 = TMA-fed persistent kernel (default), 3 = same without swizzle, 0 = cp.async 64x64 tiles, 1 = 128x64
    // updates with at most this many rows run the downdate as 64x64 tiles, 4 CTAs per SM.  Measured on B200 (profiles/
    // r01_downdate_sweep.txt) that variant wins at every k and n tried (16 resident warps hide the tile read-modify-write
    // and the operand ring better than 8 warps of one 128x128 CTA), so it is the default for all k; option 4 lowers it.
    int downdate_small_k = 1 << 30;
    int use_pdl = 1;          // programmatic dependent launch for the kernels of the frame (option 8)
    int trsm_pair = 1;        // batched filters: slab footprint that fits two CTAs per SM when possible (option 6)
    int ransac_chunk = 0;     // hypotheses evaluated per round; 0 = 16 for a single filter, 4 for batches (option 7)
    int trsm_stages = 4;      // upper limit of the slab TRSM's operand ring depth (option 5)
    int schain_variant = -1;  // -1 = automatic (single filter and k > 128: 4, else 0), 0 = one fused launch per block step
                              // (ekf_schain.cuh), 1 = panel + trail launches, 3 = the whole chain in one launch (ekf_chain.cuh),
                              // 4 = chain + slab TRSM in one launch (single filter)
    int schain_eff = 0;       // the variant the current update uses
    int* ddQueue = nullptr;   // work queues of the TMA-fed downdate: [8][2] ints (position, CTAs done), zero between launches
    int ddQueueNext = 0;      // queue the next launch uses (a lane's index; 0 without lanes)
    int dd_probe = 0;         // timing probe of the TMA-fed downdate (option 13; results are wrong when set): 1 = no DMMA, 2 = no stores, 4 = no mirror store
    int dd_ctas_per_sm = 2;   // persistent CTAs per SM of the TMA-fed downdate (option 12)
    int lanes_opt = -1;       // lanes per handle (option 11): -1 = automatic (2 for 8 or more filters), 1 = off
    std::vector<Lane> lanes;
    cudaEvent_t evLaneFork = nullptr;
    bool in_step = false;      // set by ekfb_step: phases may fold their neighbours' small launches into their own kernels
    bool mask_cleared = false; // ekfb_measure has cleared the matching mask for the ekfb_match that follows
    int small_update = 1;     // updates of at most 128 rows: factorisation + slab TRSM in one launch, the factorisation redone by
                              // every slab CTA (ekf_update_small.cuh; option 10).  Batches use it while the slab CTAs of all
                              // filters fit in two waves (beyond that the redundant factorisations cost more than the launches)
    void* tmaEncode = nullptr;   // cuTensorMapEncodeTiled (driver entry point, fetched once; no link against libcuda)
    int* chainCtl = nullptr;  // per-filter control blocks of the one-launch chain (generation, queue, flags)
    int nbMax = 0;
    bool dd_timing = false;
    std::vector<cudaEvent_t> dd_ev;   // pairs
    size_t dd_used = 0;               // events used
    double dd_flops = 0., dd_bytes = 0.;
    std::vector<double> dd_launch_flops;   // algorithmic flop of every timed launch
    bool map_ready = false;           // map-management buffers are allocated on first use
    uint8_t* mask2 = nullptr;         // new-feature mask (E/DetectNewImageFeatures.cpp:101-122), built by ekfb_map_management
    bool mask2_valid = false;
    // NCC active search (ekf_ncc.cuh): image pyramid and templates
    bool ncc_ready = false;
    uint8_t* ncc_img[kNccLevels] = {nullptr, nullptr, nullptr};
    size_t ncc_level_bytes[kNccLevels] = {0, 0, 0};
    NccView ncc;
    uint8_t* ncc_tmpl = nullptr;
    uint8_t* ncc_tmpl2 = nullptr;     // compaction targets (swapped with the live arrays by ekfb_map_management)
    double* ncc_anchor = nullptr;     // [F][Nmax][10]: camera pose + pixel at template capture (ekf_ncc.cuh)
    double* ncc_anchor2 = nullptr;
    std::vector<NccMaps> ncc_maps;    // per filter: tensor maps of its three pyramid levels (window loads of the search)
    int ncc_tma_level[kNccLevels] = {0, 0, 0};
    std::vector<uint8_t> ncc_has_image;   // per filter: a frame has been set (templates of new features are captured from it)
    int matcher = 0;                  // option 14: 0 = the reference's descriptor matcher, 1 = NCC active search inside ekfb_match
    int ncc_tma_window = 1;           // option 16: 1 = search windows by tensor-map TMA where the level allows, 0 = bulk row copies
    int ncc_warp = 1;                 // option 15: predict the template's appearance for the current camera (affine warp)
    double ncc_min_score = 0.8;       // acceptance threshold of the NCC matcher inside ekfb_match (ekfb_ncc_set_threshold)
    // device front end (ekf_frontend.cuh): corner-score image, per-row counts / offsets, keypoint count per filter
    uint8_t* fe_score = nullptr;
    uint8_t* fe_color = nullptr;   // staging for a colour frame (W x H x 4)
    int* fe_rows = nullptr;   // [F][2][H]
    int* fe_count = nullptr;  // [F]
};

template <typename T>
static int dev_alloc(ekfb_ctx* c, T** p, size_t count)
{
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    CK(cudaMemsetAsync(q, 0, std::max<size_t>(count, 1) * sizeof(T), c->stream));
    c->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return EKFB_OK;
}

#define ALLOC(ptr, count)                                   \
    do {                                                    \
        int rc_ = dev_alloc(c, &(ptr), (size_t)(count));    \
        if (rc_ != EKFB_OK) return rc_;                     \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int rup(int a, int b) { return cdiv(a, b) * b; }

struct GroupScope {
    ekfb_ctx* c;
    int g;
    GroupScope(ekfb_ctx* c_, int g_) : c(c_), g(g_)
    {
        c->cur_group = g;
        if (c->prof) cudaEventRecord(c->pe[0], c->stream);
    }
    ~GroupScope()
    {
        if (c->prof) {
            cudaEventRecord(c->pe[1], c->stream);
            cudaEventSynchronize(c->pe[1]);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->pe[0], c->pe[1]);
            c->prof_ms[g] += ms;
        }
        c->cur_group = G_MISC;
    }
};

static inline void count_launch(ekfb_ctx* c, int nlaunch = 1)
{
    c->launches += nlaunch;
    c->prof_launch[c->cur_group] += nlaunch;
}

// launch with the programmatic-stream-serialization attribute (see grid_dependency_wait in ekf_kernels.cuh);
// ctx->use_pdl = 0 (option 9) turns every such launch into a plain one
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// a frame kernel on the handle's stream: with PDL unless the handle's switch is off
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(ekfb_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args)
{
    if (c->use_pdl) return launch_pdl(kernel, grid, block, smem, c->stream, args...);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}


extern "C" const char* ekfb_last_error(void) { return g_err.c_str(); }

static int create_impl(const ekfb_params* p, int device, int n_filters, int max_features, int max_keypoints, ekfb_ctx* c);

extern "C" int ekfb_create(const ekfb_params* p, int device, int n_filters, int max_features, int max_keypoints,
                           ekfb_handle* out)
{
    REQUIRE(p && out, "null argument");
    *out = nullptr;
    REQUIRE(n_filters > 0 && max_features > 0 && max_keypoints > 0, "sizes must be positive");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        g_err = "libekf_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return EKFB_ERR_CUDA;
    }
    ekfb_ctx* c = new ekfb_ctx();
    c->smCount = prop.multiProcessorCount;
    for (int i = 0; i < 64; ++i) c->timers[i] = nullptr;
    c->pe[0] = c->pe[1] = nullptr;
    // any failure below (an out-of-memory on a large max_features is the realistic one) releases everything built so far
    const int rc = create_impl(p, device, n_filters, max_features, max_keypoints, c);
    if (rc != EKFB_OK) {
        const std::string keep = g_err;
        ekfb_destroy(c);
        cudaGetLastError();
        g_err = keep;
        return rc;
    }
    *out = c;
    return EKFB_OK;
}

static int create_impl(const ekfb_params* p, int device, int n_filters, int max_features, int max_keypoints, ekfb_ctx* c)
{
    c->prm = *p;
    c->device = device;
    c->F = n_filters;
    c->Nmax = max_features;
    c->nmax = 13 + 6 * max_features;
    c->ld = rup(c->nmax + 1, 16);
    c->Kpmax = max_keypoints;
    c->kmax = rup(2 * max_features, kNB);
    c->ldS = rup(c->kmax + 1, 16);
    c->supWords = cdiv(max_features, 32);
    std::memset(c->prof_ms, 0, sizeof(c->prof_ms));
    std::memset(c->prof_launch, 0, sizeof(c->prof_launch));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    for (int i = 0; i < 64; ++i) CK(cudaEventCreate(&c->timers[i]));
    CK(cudaEventCreate(&c->pe[0]));
    CK(cudaEventCreate(&c->pe[1]));

    const size_t F = c->F, N = c->Nmax;
    DevView& v = c->v;
    std::memset(&v, 0, sizeof(v));
    v.F = c->F; v.Nmax = c->Nmax; v.nmax = c->nmax; v.ld = c->ld; v.Kpmax = c->Kpmax; v.kmax = c->kmax;
    v.ldS = c->ldS; v.W = p->pixels_x; v.H = p->pixels_y; v.supWords = c->supWords;
    v.maxAxes = (int)(2.0 * std::max(p->pixels_x, p->pixels_y));  // E/Matching.cpp:198
    v.cam.fx = p->fx; v.cam.fy = p->fy; v.cam.k1 = p->k1; v.cam.k2 = p->k2; v.cam.cx = p->cx; v.cam.cy = p->cy;
    v.cam.dx = p->dx; v.cam.dy = p->dy; v.cam.fov_x = p->angular_vision_x; v.cam.fov_y = p->angular_vision_y;
    v.cam.width = p->pixels_x; v.cam.height = p->pixels_y;
    v.sd_lin = p->linear_accel_sd; v.sd_ang = p->angular_accel_sd; v.sigma_px = p->pixel_error_x;
    v.match_coef = p->matching_coef; v.ransac_thr = p->ransac_threshold; v.ransac_p = p->ransac_all_inliers_prob;
    v.chi2 = p->ransac_chi2;

    ALLOC(v.x, F * c->ld);
    ALLOC(v.P, F * c->nmax * c->ld);
    ALLOC(v.ftype, F * N); ALLOC(v.foff, F * N); ALLOC(v.desc, F * N * 32);
    ALLOC(v.tpred, F * N); ALLOC(v.tmatch, F * N); ALLOC(v.dims, F * D_STRIDE);
    ALLOC(v.vis, F * N); ALLOC(v.h, F * N * 2); ALLOC(v.Si, F * N * 4); ALLOC(v.Hx, F * N * 14); ALLOC(v.Hf, F * N * 12);
    ALLOC(v.ellax, F * N * 2); ALLOC(v.ellang, F * N);
    ALLOC(v.vis2, F * N); ALLOC(v.h2, F * N * 2); ALLOC(v.Si2, F * N * 4); ALLOC(v.Hx2, F * N * 14); ALLOC(v.Hf2, F * N * 12);
    ALLOC(v.mflag, F * N); ALLOC(v.z, F * N * 2); ALLOC(v.mkp, F * N); ALLOC(v.mdist, F * N); ALLOC(v.mlist, F * N);
    ALLOC(v.inl, F * N); ALLOC(v.outl, F * N); ALLOC(v.resc, F * N); ALLOC(v.ulist, F * N);
    ALLOC(v.kpok, F * c->Kpmax); ALLOC(v.mask, F * (size_t)v.W * v.H);
    ALLOC(v.hypcount, F * N); ALLOC(v.hypsup, F * N * c->supWords);
    ALLOC(v.Bu, F * c->kmax * c->ld); ALLOC(v.S, F * c->kmax * c->ldS); ALLOC(v.Sf, F * c->kmax * c->ldS); ALLOC(v.dx, F * c->ld); ALLOC(v.dbg, 64); ALLOC(v.Uinv, F * (c->kmax / kNB) * kNB * kNB); ALLOC(v.Jq, F * 16);
    ALLOC(c->d_kpxy, F * c->Kpmax * 2); ALLOC(c->d_kpdesc, F * c->Kpmax * 32);
    ALLOC(c->d_kpxy_ptr, F); ALLOC(c->d_kpdesc_ptr, F); ALLOC(c->d_kpcount, F);
    ALLOC(c->d_rec, F);
    ALLOC(c->ddQueue, 16);
    c->nbMax = c->kmax / kNB + 1;
    ALLOC(c->chainCtl, F * (size_t)chain_ctl_ints(c->nbMax));
    v.kpxy = c->d_kpxy_ptr;
    v.kpdesc = c->d_kpdesc_ptr;
    CK(cudaMallocHost(&c->h_dims, F * D_STRIDE * sizeof(int)));
    {
        int* zc = nullptr;
        CK(cudaHostAlloc(&zc, (F * D_STRIDE + F) * sizeof(int), cudaHostAllocMapped));
        std::memset(zc, 0, (F * D_STRIDE + F) * sizeof(int));
        c->h_dims_zc = zc;
        c->h_flag_zc = zc + F * D_STRIDE;
        int* dzc = nullptr;
        CK(cudaHostGetDevicePointer(&dzc, zc, 0));
        v.hostDims = dzc;
        v.hostFlag = dzc + F * D_STRIDE;
    }
    CK(cudaMallocHost(&c->h_kpxy_ptr, F * sizeof(void*)));
    CK(cudaMallocHost(&c->h_kpdesc_ptr, F * sizeof(void*)));
    CK(cudaMallocHost(&c->h_kpcount, F * sizeof(int)));
    CK(cudaMallocHost(&c->h_rec, F * sizeof(RecordDev)));
    std::memset(c->h_dims, 0, F * D_STRIDE * sizeof(int));
    for (int f = 0; f < c->F; ++f) {
        c->h_kpxy_ptr[f] = c->d_kpxy + (size_t)f * c->Kpmax * 2;
        c->h_kpdesc_ptr[f] = c->d_kpdesc + (size_t)f * c->Kpmax * 32;
    }
    CK(cudaMemcpyAsync(c->d_kpxy_ptr, c->h_kpxy_ptr, F * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_kpdesc_ptr, c->h_kpdesc_ptr, F * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    c->hn.assign(c->F, 13);
    c->hN.assign(c->F, 0);
    c->hKp.assign(c->F, 0);
    c->seq.resize(c->F);

    CK(cudaFuncSetAttribute(k_gemm_tn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    CK(cudaFuncSetAttribute(k_gemm_tn<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    CK(cudaFuncSetAttribute(k_gemm_tn<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    CK(cudaFuncSetAttribute(k_downdate, cudaFuncAttributeMaxDynamicSharedMemorySize, kDownSmemBytes));
    CK(cudaFuncSetAttribute(k_downdate64, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
    CK(cudaFuncSetAttribute(k_schain_trail, cudaFuncAttributeMaxDynamicSharedMemorySize, kSTrailSmem));
    CK(cudaFuncSetAttribute(k_schain_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSPanelSmem));
    CK(cudaFuncSetAttribute(k_schain_step, cudaFuncAttributeMaxDynamicSharedMemorySize, kStepSmem));
    CK(cudaFuncSetAttribute(k_schain_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem));
    CK(cudaFuncSetAttribute(k_downdate_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTdSmemBytes));
    CK(cudaFuncSetAttribute(k_downdate_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTdSmemBytes));
    {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->tmaEncode, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            c->tmaEncode = nullptr;
        cudaGetLastError();
    }
    CK(cudaFuncSetAttribute(k_update_fused<24, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_update_fused<24, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_update_fused<24, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_update_fused<24, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_gemm_tn_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kGtSmemBytes));
    CK(cudaFuncSetAttribute(k_update_small<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_small_smem_bytes(24)));
    CK(cudaFuncSetAttribute(k_trsm_slab<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_trsm_slab<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_trsm_slab<24, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_trsm_slab<24, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_trsm_slab<24, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    CK(cudaFuncSetAttribute(k_ransac_hyp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            std::min<int>(227 * 1024 - 2048, c->ld * (int)sizeof(double))));
    const int rasterSmem = 4 * (int)(sizeof(RasterScratch) + sizeof(int) * 2 * (size_t)v.H);
    CK(cudaFuncSetAttribute(k_mask_raster, cudaFuncAttributeMaxDynamicSharedMemorySize, rasterSmem));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_destroy(ekfb_handle c)
{
    if (!c) return EKFB_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void* p : c->allocs) cudaFree(p);
    for (Lane& L : c->lanes) {
        if (L.stream && L.stream != c->stream) cudaStreamDestroy(L.stream);
        if (L.done) cudaEventDestroy(L.done);
    }
    if (c->evLaneFork) cudaEventDestroy(c->evLaneFork);
    for (SeqDev& s : c->seq) {
        if (s.xy) cudaFree(s.xy);
        if (s.desc) cudaFree(s.desc);
    }
    if (c->flush_buf) cudaFree(c->flush_buf);
    for (cudaEvent_t e : c->dd_ev) cudaEventDestroy(e);
    if (c->h_dims) cudaFreeHost(c->h_dims);
    if (c->h_dims_zc) cudaFreeHost(c->h_dims_zc);
    if (c->h_kpxy_ptr) cudaFreeHost(c->h_kpxy_ptr);
    if (c->h_kpdesc_ptr) cudaFreeHost(c->h_kpdesc_ptr);
    if (c->h_kpcount) cudaFreeHost(c->h_kpcount);
    if (c->h_rec) cudaFreeHost(c->h_rec);
    for (int i = 0; i < 64; ++i)
        if (c->timers[i]) cudaEventDestroy(c->timers[i]);
    if (c->pe[0]) cudaEventDestroy(c->pe[0]);
    if (c->pe[1]) cudaEventDestroy(c->pe[1]);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    delete c;
    return EKFB_OK;
}

extern "C" int ekfb_sync(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

// ---- state in / out ---------------------------------------------------------------------------
extern "C" int ekfb_set_state(ekfb_handle c, int f, int n, int N, const double* x, const int32_t* ftype,
                              const int32_t* foff, const double* P, const uint8_t* desc)
{
    REQUIRE(c && x && P && (N == 0 || (ftype && foff)), "null argument");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    if (n > c->nmax || N > c->Nmax || n < 13) {
        g_err = "state exceeds the capacity reserved by ekfb_create";
        return EKFB_ERR_CAPACITY;
    }
    CK(cudaSetDevice(c->device));
    DevView& v = c->v;
    CK(cudaMemsetAsync(v.x + (size_t)f * c->ld, 0, sizeof(double) * c->ld, c->stream));
    CK(cudaMemcpyAsync(v.x + (size_t)f * c->ld, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpy2DAsync(v.P + (size_t)f * c->nmax * c->ld, sizeof(double) * c->ld, P, sizeof(double) * n,
                         sizeof(double) * n, n, cudaMemcpyHostToDevice, c->stream));
    if (N > 0) {
        CK(cudaMemcpyAsync(v.ftype + (size_t)f * c->Nmax, ftype, sizeof(int) * N, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(v.foff + (size_t)f * c->Nmax, foff, sizeof(int) * N, cudaMemcpyHostToDevice, c->stream));
        if (desc)
            CK(cudaMemcpyAsync(v.desc + (size_t)f * c->Nmax * 32, desc, (size_t)N * 32, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(v.tpred + (size_t)f * c->Nmax, 0, sizeof(int) * N, c->stream));
        CK(cudaMemsetAsync(v.tmatch + (size_t)f * c->Nmax, 0, sizeof(int) * N, c->stream));
    }
    int* hd = c->h_dims + (size_t)f * D_STRIDE;
    std::memset(hd, 0, sizeof(int) * D_STRIDE);
    hd[D_N_STATE] = n;
    hd[D_N_FEAT] = N;
    hd[D_N_KP] = c->hKp[f];
    CK(cudaMemcpyAsync(v.dims + (size_t)f * D_STRIDE, hd, sizeof(int) * D_STRIDE, cudaMemcpyHostToDevice, c->stream));
    k_symmetrize<<<dim3(cdiv(n, 16), cdiv(n, 16)), dim3(16, 16), 0, c->stream>>>(v, f);
    count_launch(c);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));  // caller buffers may be pageable
    c->hn[f] = n;
    c->hN[f] = N;
    return EKFB_OK;
}

extern "C" int ekfb_get_state(ekfb_handle c, int f, double* x, double* P, int cam_only)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const int n = cam_only ? 13 : c->hn[f];
    if (x) CK(cudaMemcpyAsync(x, c->v.x + (size_t)f * c->ld, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (P)
        CK(cudaMemcpy2DAsync(P, sizeof(double) * n, c->v.P + (size_t)f * c->nmax * c->ld, sizeof(double) * c->ld,
                             sizeof(double) * n, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_get_descriptors(ekfb_handle c, int f, uint8_t* desc, int32_t* tp, int32_t* tm)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const int N = c->hN[f];
    if (desc) CK(cudaMemcpyAsync(desc, c->v.desc + (size_t)f * c->Nmax * 32, (size_t)N * 32, cudaMemcpyDeviceToHost, c->stream));
    if (tp) CK(cudaMemcpyAsync(tp, c->v.tpred + (size_t)f * c->Nmax, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
    if (tm) CK(cudaMemcpyAsync(tm, c->v.tmatch + (size_t)f * c->Nmax, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_get_dims(ekfb_handle c, int f, int32_t* n, int32_t* N)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    if (n) *n = c->hn[f];
    if (N) *N = c->hN[f];
    return EKFB_OK;
}

// ---- keypoints ----------------------------------------------------------------------------------
static int push_kp_meta(ekfb_ctx* c)
{
    // pointer tables and counts of all filters: three copies and one launch, whatever the number of filters
    const size_t F = c->F;
    for (int f = 0; f < c->F; ++f) c->h_dims[(size_t)f * D_STRIDE + D_N_KP] = c->h_kpcount[f] = c->hKp[f];
    CK(cudaMemcpyAsync(c->d_kpxy_ptr, c->h_kpxy_ptr, F * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_kpdesc_ptr, c->h_kpdesc_ptr, F * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_kpcount, c->h_kpcount, F * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_set_kp_counts<<<cdiv(c->F, 128), 128, 0, c->stream>>>(c->v, c->d_kpcount);
    count_launch(c);
    CK(cudaGetLastError());
    return EKFB_OK;
}

extern "C" int ekfb_set_keypoints(ekfb_handle c, int f, const float* xy, const uint8_t* desc, int n_kp)
{
    REQUIRE(c && (n_kp == 0 || (xy && desc)), "null argument");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    if (n_kp > c->Kpmax || n_kp < 0) {
        g_err = "keypoint count exceeds the capacity reserved by ekfb_create";
        return EKFB_ERR_CAPACITY;
    }
    CK(cudaSetDevice(c->device));
    float* dxy = c->d_kpxy + (size_t)f * c->Kpmax * 2;
    uint8_t* dds = c->d_kpdesc + (size_t)f * c->Kpmax * 32;
    if (n_kp > 0) {
        CK(cudaMemcpyAsync(dxy, xy, sizeof(float) * 2 * n_kp, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(dds, desc, (size_t)32 * n_kp, cudaMemcpyHostToDevice, c->stream));
    }
    c->hKp[f] = n_kp;
    c->h_kpxy_ptr[f] = dxy;
    c->h_kpdesc_ptr[f] = dds;
    c->h_dims[(size_t)f * D_STRIDE + D_N_KP] = n_kp;
    CK(cudaMemcpyAsync(c->d_kpxy_ptr + f, c->h_kpxy_ptr + f, sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_kpdesc_ptr + f, c->h_kpdesc_ptr + f, sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v.dims + (size_t)f * D_STRIDE + D_N_KP, c->h_dims + (size_t)f * D_STRIDE + D_N_KP, sizeof(int),
                       cudaMemcpyHostToDevice, c->stream));
    return EKFB_OK;
}

// ekfb_set_keypoints for every filter of the handle in one call: xy[f] / desc[f] = host buffers of filter f, n_kp[f] its count
extern "C" int ekfb_set_keypoints_batch(ekfb_handle c, const float* const* xy, const uint8_t* const* desc, const int32_t* n_kp)
{
    REQUIRE(c && xy && desc && n_kp, "null argument");
    CK(cudaSetDevice(c->device));
    for (int f = 0; f < c->F; ++f) {
        if (n_kp[f] > c->Kpmax || n_kp[f] < 0) {
            g_err = "keypoint count exceeds the capacity reserved by ekfb_create";
            return EKFB_ERR_CAPACITY;
        }
        REQUIRE(n_kp[f] == 0 || (xy[f] && desc[f]), "null keypoint buffer");
    }
    for (int f = 0; f < c->F; ++f) {
        float* dxy = c->d_kpxy + (size_t)f * c->Kpmax * 2;
        uint8_t* dds = c->d_kpdesc + (size_t)f * c->Kpmax * 32;
        if (n_kp[f] > 0) {
            CK(cudaMemcpyAsync(dxy, xy[f], sizeof(float) * 2 * n_kp[f], cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(dds, desc[f], (size_t)32 * n_kp[f], cudaMemcpyHostToDevice, c->stream));
        }
        c->hKp[f] = n_kp[f];
        c->h_kpxy_ptr[f] = dxy;
        c->h_kpdesc_ptr[f] = dds;
    }
    return push_kp_meta(c);
}

extern "C" int ekfb_set_keypoints_packed(ekfb_handle c, const float* xy, const uint8_t* desc, const int32_t* kp_offset)
{
    REQUIRE(c && kp_offset, "null argument");
    CK(cudaSetDevice(c->device));
    const int total = kp_offset[c->F];
    REQUIRE(kp_offset[0] == 0 && (total == 0 || (xy && desc)), "bad offsets or null keypoint buffer");
    for (int f = 0; f < c->F; ++f) {
        const int n = kp_offset[f + 1] - kp_offset[f];
        if (n < 0 || n > c->Kpmax) {
            g_err = "keypoint count exceeds the capacity reserved by ekfb_create";
            return EKFB_ERR_CAPACITY;
        }
    }
    // the staging area holds F * Kpmax keypoints, so the packed frame always fits; filter f reads it at its offset
    if (total > 0) {
        CK(cudaMemcpyAsync(c->d_kpxy, xy, sizeof(float) * 2 * (size_t)total, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_kpdesc, desc, (size_t)32 * total, cudaMemcpyHostToDevice, c->stream));
    }
    for (int f = 0; f < c->F; ++f) {
        c->hKp[f] = kp_offset[f + 1] - kp_offset[f];
        c->h_kpxy_ptr[f] = c->d_kpxy + (size_t)kp_offset[f] * 2;
        c->h_kpdesc_ptr[f] = c->d_kpdesc + (size_t)kp_offset[f] * 32;
    }
    return push_kp_meta(c);
}

extern "C" int ekfb_load_sequence(ekfb_handle c, int f, int n_frames, const int32_t* kp_offset, const float* xy,
                                  const uint8_t* desc)
{
    REQUIRE(c && kp_offset && xy && desc && n_frames > 0, "null argument");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    SeqDev& s = c->seq[f];
    if (s.xy) cudaFree(s.xy);
    if (s.desc) cudaFree(s.desc);
    s.xy = nullptr;
    s.desc = nullptr;
    s.off.assign(kp_offset, kp_offset + n_frames + 1);
    for (int t = 0; t < n_frames; ++t)
        if (s.off[t + 1] - s.off[t] > c->Kpmax || s.off[t + 1] < s.off[t]) {
            g_err = "a frame of the sequence exceeds max_keypoints";
            return EKFB_ERR_CAPACITY;
        }
    const size_t total = (size_t)s.off[n_frames];
    CK(cudaMalloc(&s.xy, std::max<size_t>(total, 1) * 2 * sizeof(float)));
    CK(cudaMalloc(&s.desc, std::max<size_t>(total, 1) * 32));
    CK(cudaMemcpyAsync(s.xy, xy, total * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(s.desc, desc, total * 32, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_select_frame(ekfb_handle c, int frame)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    for (int f = 0; f < c->F; ++f) {
        SeqDev& s = c->seq[f];
        REQUIRE(s.xy && frame >= 0 && frame + 1 < (int)s.off.size(), "no such frame in the loaded sequence");
        c->h_kpxy_ptr[f] = s.xy + (size_t)s.off[frame] * 2;
        c->h_kpdesc_ptr[f] = s.desc + (size_t)s.off[frame] * 32;
        c->hKp[f] = s.off[frame + 1] - s.off[frame];
    }
    return push_kp_meta(c);
}

// ---- phases ----------------------------------------------------------------------------------------
static int max_of(const std::vector<int>& a)
{
    int m = 0;
    for (int x : a) m = std::max(m, x);
    return m;
}

static int read_dims(ekfb_ctx* c)
{
    CK(cudaMemcpyAsync(c->h_dims, c->v.dims, (size_t)c->F * D_STRIDE * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

// Wait for the counters a publishing kernel (publish_dims) launched with sequence number `seq` wrote into mapped host
// memory.  Falls back to the copy + synchronise path if the stream drains without the flags (a failed launch).
static int wait_published_dims(ekfb_ctx* c, int seq)
{
    const int F = c->F;
    for (unsigned spin = 1;; ++spin) {
        bool all = true;
        for (int f = 0; f < F; ++f) all = all && c->h_flag_zc[f] == seq;
        if (all) break;
        if ((spin & 0x3ff) == 0) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaErrorNotReady) continue;
            CK(q);
            all = true;
            for (int f = 0; f < F; ++f) all = all && c->h_flag_zc[f] == seq;
            if (all) break;
            return read_dims(c);
        }
    }
    std::memcpy(c->h_dims, c->h_dims_zc, (size_t)F * D_STRIDE * sizeof(int));
    return EKFB_OK;
}

extern "C" int ekfb_predict(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    GroupScope gs(c, G_PREDICT);
    const int n = max_of(c->hn);
    dim3 grid(1 + cdiv(std::max(n - 13, 0), 256), c->F);
    CK(launch_k(c, k_predict_cov, grid, dim3(256), 0, c->v));   // + the state prediction, by the block that finishes last
    count_launch(c);
    CK(cudaGetLastError());
    return EKFB_OK;
}

static int launch_measure(ekfb_ctx* c, int mode, int seq = 0, int extra = 0)
{
    const int N = std::max(max_of(c->hN), mode ? 1 : 0);   // (the rescue pass must run its tail even for an empty map)
    if (N == 0) return EKFB_OK;
    dim3 grid(cdiv(N, 8), c->F);
    CK(launch_k(c, k_measure, grid, dim3(256), 0, c->v, mode, seq, extra));
    count_launch(c);
    CK(cudaGetLastError());
    return EKFB_OK;
}

extern "C" int ekfb_measure(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    GroupScope gs(c, G_MEASURE);
    // inside ekfb_step the measurement kernel also clears the matching mask (one launch less in front of the rasteriser)
    const bool clear = c->in_step && ((size_t)c->v.W * c->v.H) % 16 == 0 && max_of(c->hN) > 0;
    c->mask_cleared = clear;
    return launch_measure(c, 0, 0, clear ? 1 : 0);
}

extern "C" int ekfb_match(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    if (c->matcher == 1 && c->ncc_ready) {   // EKFB_OPT_MATCHER: the NCC active search in place of the descriptor matcher
        c->mask_cleared = false;
        return ekfb_match_ncc(c, c->ncc_min_score);
    }
    GroupScope gs(c, G_MATCH);
    const int N = max_of(c->hN), Kp = max_of(c->hKp);
    DevView& v = c->v;
    if (!c->mask_cleared) CK(cudaMemsetAsync(v.mask, 0, (size_t)c->F * v.W * v.H, c->stream));
    c->mask_cleared = false;
    if (N > 0) {
        const int rasterSmem = 4 * (int)(sizeof(RasterScratch) + sizeof(int) * 2 * (size_t)v.H);
        CK(launch_k(c, k_mask_raster, dim3(cdiv(N, 4), c->F), dim3(128), rasterSmem, v, v.mask, v.maxAxes, 255));
        count_launch(c);
        // the detector-mask test of the keypoints and the bookkeeping after matching (counts, match list, RANSAC reset) are
        // part of the matching kernel: staged with the keypoints / run by the block that finishes last
        CK(launch_k(c, k_match, dim3(cdiv(N, 8), c->F), dim3(256), 0, v, 1, 1));
        count_launch(c);
    } else {
        CK(launch_k(c, k_after_match, dim3(c->F), dim3(256), 0, v));
        count_launch(c);
    }
    CK(cudaGetLastError());
    return EKFB_OK;
}

// hypotheses per round: the adaptive cap of the reference's rule collapses to 2-5 once a majority hypothesis is seen, so
// batches (where surplus hypotheses cost real throughput) go four at a time, a single filter (latency) sixteen
static int ransac_chunk_len(const ekfb_ctx* c) { return c->ransac_chunk > 0 ? std::min(c->ransac_chunk, 64) : (c->F >= 8 ? 4 : 16); }

// one round: hypotheses chunk0 .. chunk0 + CH - 1 of every filter, then the sequential acceptance rule; *seq = the sequence
// number the counters will be published under
static int ransac_launch(ekfb_ctx* c, int chunk0, int* seq)
{
    const int CH = ransac_chunk_len(c);
    *seq = ++c->dims_seq;
    // (the acceptance rule + the inlier split are run by the hypothesis block that finishes last: one launch per round)
    CK(launch_k(c, k_ransac_hyp, dim3(CH, c->F, cdiv(c->Nmax, kHypFeat)), dim3(416), 0, c->v, chunk0, *seq));
    count_launch(c);
    CK(cudaGetLastError());
    return EKFB_OK;
}

// waits for the counters of the round; *done = every filter has finished (or no hypotheses are left)
static int ransac_collect(ekfb_ctx* c, int seq, int chunk0, bool* done)
{
    int rc = wait_published_dims(c, seq);
    if (rc != EKFB_OK) return rc;
    bool all = true;
    for (int f = 0; f < c->F; ++f) all = all && c->h_dims[(size_t)f * D_STRIDE + D_RANSAC_DONE];
    *done = all || chunk0 + ransac_chunk_len(c) >= std::max(max_of(c->hN), 1);
    return EKFB_OK;
}

extern "C" int ekfb_ransac(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    GroupScope gs(c, G_RANSAC);
    const int CH = ransac_chunk_len(c);
    for (int chunk0 = 0;; chunk0 += CH) {
        int seq = 0;
        bool done = false;
        int rc = ransac_launch(c, chunk0, &seq);
        if (rc == EKFB_OK) rc = ransac_collect(c, seq, chunk0, &done);
        if (rc != EKFB_OK) return rc;
        if (done) break;
    }
    return EKFB_OK;
}

// P -= W W^T for all filters of the handle (W^T = rows of Bu, K = 2 * ulist count per filter)
static int launch_downdate(ekfb_ctx* c, int n, bool allowTma = true)
{
    DevView& v = c->v;
    const int nI = cdiv(n, kDTM);
    int kMax = 0;
    for (int f = 0; f < c->F; ++f) kMax = std::max(kMax, 2 * c->h_dims[(size_t)f * D_STRIDE + D_ULIST]);
    const bool timeIt = c->dd_timing && c->dd_used + 2 <= c->dd_ev.size();
    if (timeIt) cudaEventRecord(c->dd_ev[c->dd_used], c->stream);
    if (c->downdate_variant >= 2 && allowTma && c->tmaEncode && kMax > 0) {
        // TMA-fed persistent kernel (ekf_downdate_tma.cuh): 3-D tensor maps over (column, row, filter) of P and of W^T, encoded
        // per launch (the P buffers swap under map management).  Needs the rows of W^T between K and the end of its last 16-row
        // chunk to be zero, which the slab TRSM guarantees (allowTma is false after the generic factorisation).
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        const bool swz = c->downdate_variant == 2;
        TdMaps maps;
        const cuuint64_t pd[3] = {(cuuint64_t)c->nmax, (cuuint64_t)c->nmax, (cuuint64_t)c->F};
        const cuuint64_t wd[3] = {(cuuint64_t)c->ld, (cuuint64_t)c->kmax, (cuuint64_t)c->F};
        const cuuint64_t pst[2] = {(cuuint64_t)c->ld * sizeof(double), (cuuint64_t)c->nmax * c->ld * sizeof(double)};
        const cuuint64_t wst[2] = {(cuuint64_t)c->ld * sizeof(double), (cuuint64_t)c->kmax * c->ld * sizeof(double)};
        const cuuint32_t pb[3] = {swz ? 16u : 64u, 64u, 1u}, wb[3] = {swz ? 16u : 64u, 16u, 1u}, es[3] = {1, 1, 1};
        const CUtensorMapSwizzle sw = swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
        CUresult r1 = ((EncodeFn)c->tmaEncode)(&maps.P, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, v.P, pd, pst, pb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = ((EncodeFn)c->tmaEncode)(&maps.W, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, v.Bu, wd, wst, wb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
            g_err = "cuTensorMapEncodeTiled failed for the downdate operands";
            return EKFB_ERR_CUDA;
        }
        const int nT = cdiv(n, 64), tilesMax = nT * (nT + 1) / 2;
        const dim3 grid((unsigned)std::min<long long>((long long)tilesMax * c->F, (long long)c->dd_ctas_per_sm * c->smCount));
        int* queue = c->ddQueue + 2 * c->ddQueueNext;   // the handle's lanes (streams) launch concurrently: a queue each
        if (swz) CK(launch_k(c, k_downdate_tma<true>, grid, dim3(160), (size_t)kTdSmemBytes, v, maps, tilesMax, nT, queue, c->dd_probe));
        else CK(launch_k(c, k_downdate_tma<false>, grid, dim3(160), (size_t)kTdSmemBytes, v, maps, tilesMax, nT, queue, c->dd_probe));
    } else if (c->downdate_variant == 1)
        k_downdate<<<dim3(nI * (nI + 1), c->F), 128, kDownSmemBytes, c->stream>>>(v);
    else if (kMax <= c->downdate_small_k) {
        // 64x64 tiles, four CTAs per SM (the default at every k, see downdate_small_k)
        CK(launch_k(c, k_downdate64, dim3(nI * (nI + 1) / 2 * 4, c->F), dim3(128), kSmallSmemBytes, v, 0));
    } else {
        // 1-D grid over the T lower 128x128 tiles.  Single filter: if T is just above a multiple of the SM
        // count, the remainder would cost a whole extra wave; it runs as 64x64 tiles on a second stream,
        // co-resident with the big CTAs (register / smem budgets of the two kernels are sized for that).
        const int T = nI * (nI + 1) / 2;
        int rem = (c->F == 1 && T > c->smCount) ? T % c->smCount : 0;
        if (rem > c->smCount / 4) rem = 0;
        const int nBig = T - rem;
        if (rem > 0) {
            CK(cudaEventRecord(c->evFork, c->stream));
            CK(cudaStreamWaitEvent(c->stream2, c->evFork, 0));
            k_downdate64<<<dim3(rem * 4, c->F), 128, kSmallSmemBytes, c->stream2>>>(v, nBig);
            CK(cudaEventRecord(c->evJoin, c->stream2));
            count_launch(c);
        }
        if (c->F == 1)
            k_gemm_tn<2><<<dim3(nBig, 1, c->F), 256, kGemmSmemBytes, c->stream>>>(v, nBig);
        else
            k_gemm_tn<3><<<dim3(nBig, 1, c->F), 256, kGemmSmemBytes, c->stream>>>(v, nBig);
        if (rem > 0) CK(cudaStreamWaitEvent(c->stream, c->evJoin, 0));
    }
    if (timeIt) {
        cudaEventRecord(c->dd_ev[c->dd_used + 1], c->stream);
        c->dd_used += 2;
        double lf = 0.;
        for (int f = 0; f < c->F; ++f) {
            const double nf = c->hn[f], kf = 2.0 * c->h_dims[(size_t)f * D_STRIDE + D_ULIST];
            lf += nf * (nf + 1.0) * kf;
        }
        c->dd_launch_flops.push_back(lf);
        for (int f = 0; f < c->F; ++f) {
            const double nf = c->hn[f], kf = 2.0 * c->h_dims[(size_t)f * D_STRIDE + D_ULIST];
            if (kf > 0) {
                c->dd_flops += nf * (nf + 1.0) * kf;   // symmetric rank-k form
                c->dd_bytes += 16.0 * nf * nf;         // read + write P once
            }
        }
    }
    return EKFB_OK;
}

// Cholesky of [S | nu] for all filters of the handle (k = the largest 2 * ulist count): factor rows into Sf, inverses of
// the diagonal blocks into Uinv
static int launch_schain(ekfb_ctx* c, int k)
{
    DevView& v = c->v;
    const int steps = cdiv(k, kNB);
    if (c->schain_eff == 1) {
        for (int J = 0; J < steps; ++J) {
            const int J0 = J * kNB, J1 = J0 + kNB;
            const int Jr = std::min(J1, k);
            const int pcols = (c->F >= 8) ? 128 : kSPanelCols;
            k_schain_panel<<<dim3(cdiv(k + 1 - Jr, pcols), c->F), 128, kSPanelSmem, c->stream>>>(v, J, pcols);
            count_launch(c);
            if (k > J1) {
                k_schain_trail<<<dim3(cdiv(k + 1 - J1, 64), cdiv(k - J1, 64), c->F), 128, kSTrailSmem, c->stream>>>(v, J);
                count_launch(c);
            }
        }
        CK(cudaMemcpyAsync(v.Sf, v.S, sizeof(double) * (size_t)c->F * c->kmax * c->ldS, cudaMemcpyDeviceToDevice, c->stream));
    } else if (c->schain_eff >= 3) {
        // the whole chain in one launch: G CTAs per filter (one critical CTA + queue workers).  Grid sized from the handle's
        // capacity, not from this frame's k: surplus CTAs find the queue empty and exit.
        const int nbR = c->kmax / kNB, nbC = nbR + 1;
        const int tasks = nbR * (nbC - 1) - nbR * (nbR - 1) / 2 + std::max(nbR - 2, 0);
        const int G = (c->F == 1) ? std::min(c->smCount, tasks + 1) : std::max(1, std::min(tasks + 1, c->smCount / c->F));
        CK(launch_pdl(k_schain_fused, dim3(G, c->F), dim3(256), (size_t)kChainSmem, c->stream, v, c->chainCtl, c->nbMax));
        count_launch(c);
    } else {
        // one launch per block step; step J needs a launch while a tile row J+1 (or the lone nu column tile) exists
        const int nbR = steps, nbC = (k + kNB) / kNB;
        for (int J = -1; J + 1 < nbC && J + 1 <= nbR; ++J) {
            if (J + 1 == nbR && nbC == nbR) break;
            CK(launch_pdl(k_schain_step, dim3(J < 0 ? 1 : nbC - (J + 1), J < 0 ? 1 : std::max(nbR - (J + 1), 1), c->F), dim3(256),
                          (size_t)kStepSmem, c->stream, v, J));
            count_launch(c);
        }
    }
    return EKFB_OK;
}

// update() for the list currently in ulist (host mirror of the counts must be fresh)
static int run_update(ekfb_ctx* c, int which)
{
    DevView& v = c->v;
    int ku = 0;
    for (int f = 0; f < c->F; ++f) ku = std::max(ku, c->h_dims[(size_t)f * D_STRIDE + D_ULIST]);
    if (ku == 0) return EKFB_OK;
    const int k = 2 * ku, n = max_of(c->hn);
    bool usedGeneric = false;
    {
        GroupScope gs(c, G_GAIN);
        CK(launch_k(c, k_gain_rows, dim3(cdiv(c->ld, 256), ku, c->F), dim3(256), 0, v, which));
        CK(launch_k(c, k_build_S, dim3(cdiv(k, 32), cdiv(k, 32), c->F), dim3(32, 8), 0, v, which));
        count_launch(c, 2);
    }
    {
        GroupScope gs(c, G_CHOL);
        const int steps = cdiv(k, kNB);
        const size_t smem16 = trsm_smem_bytes(k, 16);
        const size_t smemMax = kFusedSmemMax;   // (the kernels also hold a few bytes of static shared memory)
        // single filter, variant 4: chain and slab TRSM overlapped in one launch (ekf_chain.cuh), when one slab per SM fits
        // beside the chain's CTAs; otherwise the chain (one launch) followed by the slab TRSM
        const int fusedSlabs = cdiv(n, 24);
        // operand ring of the TRSM role as deep as the shared memory beside the slab allows: 3 stages up to k = 640, 2 up to 704,
        // a single buffer (load, then use) up to k = 832 -- slower TRSM, but it still runs under the chain instead of behind it
        // ... and with a dense slab pitch (bank conflicts on the operand loads) up to k = 1024
        int fusedNS = trsm_smem_bytes(k, 24, 3) <= (size_t)kFusedSmemMax ? 3 : (trsm_smem_bytes(k, 24, 2) <= (size_t)kFusedSmemMax ? 2 :
                      (trsm_smem_bytes(k, 24, 1) <= (size_t)kFusedSmemMax ? 1 : 0));
        int fusedPad = 4;
        if (fusedNS == 0 && trsm_smem_bytes(k, 24, 1, 0) <= (size_t)kFusedSmemMax) { fusedNS = 1; fusedPad = 0; }
        const bool fusedOk = c->F == 1 && !c->force_generic && fusedNS > 0 && fusedSlabs + 1 + 8 <= c->smCount;
        // automatic choice (measured on B200, profiles/r02_chain_downdate_variants.txt): the fused launch wins once the chain has
        // three or more block steps; below that (and for batches) one launch per block step is faster
        c->schain_eff = c->schain_variant >= 0 ? c->schain_variant : ((fusedOk && k > 128) ? 4 : 0);
        const bool smallOk = c->small_update && c->schain_variant < 0 && !c->force_generic && k <= 2 * kNB &&
                             (long long)c->F * cdiv(n, 24) <= 2ll * c->smCount;
        if (smallOk) {
            // small update: every slab CTA factors S itself, then solves its slab (ekf_update_small.cuh)
            CK(launch_pdl(k_update_small<24>, dim3(cdiv(n, 24), c->F), dim3(256), update_small_smem_bytes(24), c->stream, v));
            count_launch(c);
        } else if (c->schain_eff == 4 && fusedOk) {
            const size_t sm = std::max<size_t>(kChainSmem, trsm_smem_bytes(k, 24, fusedNS, fusedPad));
            if (fusedNS == 3)
                CK(launch_pdl(k_update_fused<24, 3>, dim3(c->smCount), dim3(256), sm, c->stream, v, c->chainCtl, c->nbMax, fusedSlabs));
            else if (fusedNS == 2)
                CK(launch_pdl(k_update_fused<24, 2>, dim3(c->smCount), dim3(256), sm, c->stream, v, c->chainCtl, c->nbMax, fusedSlabs));
            else if (fusedPad == 4)
                CK(launch_pdl(k_update_fused<24, 1>, dim3(c->smCount), dim3(256), sm, c->stream, v, c->chainCtl, c->nbMax, fusedSlabs));
            else
                CK(launch_pdl(k_update_fused<24, 1, 0>, dim3(c->smCount), dim3(256), sm, c->stream, v, c->chainCtl, c->nbMax, fusedSlabs));
            count_launch(c);   // (its last block also advances the chain's generation and applies the state correction)
        } else if (smem16 <= smemMax && !c->force_generic && (k <= c->slab_trsm_max_k || c->F > 1 || !c->tmaEncode)) {
            // fast path: S-only chain, diagonal-block inverses, slab TRSM
            { int rcC = launch_schain(c, k); if (rcC != EKFB_OK) return rcC; }
            // widest slab that fits, then the deepest operand ring beside it (4, 3 or 2 stages of 32 rows)
            const int maxStages = c->trsm_stages;
            // batched filters: many slabs per SM, so prefer a footprint that lets two CTAs share an SM (their barrier
            // stalls overlap) over the widest slab
            if (c->F > 1 && c->trsm_pair && 2 * trsm_smem_bytes(k, 24, 2) <= smemMax)
                CK(launch_pdl(k_trsm_slab<24, 2>, dim3(cdiv(n, 24), c->F), dim3(256), trsm_smem_bytes(k, 24, 2), c->stream, v));
            else if (c->F > 1 && c->trsm_pair && 2 * trsm_smem_bytes(k, 16, 2) <= smemMax)
                CK(launch_pdl(k_trsm_slab<16, 2>, dim3(cdiv(n, 16), c->F), dim3(256), trsm_smem_bytes(k, 16, 2), c->stream, v));
            else if (trsm_smem_bytes(k, 24, 4) <= smemMax && maxStages >= 4)
                CK(launch_pdl(k_trsm_slab<24, 4>, dim3(cdiv(n, 24), c->F), dim3(256), trsm_smem_bytes(k, 24, 4), c->stream, v));
            else if (trsm_smem_bytes(k, 24, 3) <= smemMax && maxStages >= 3)
                CK(launch_pdl(k_trsm_slab<24, 3>, dim3(cdiv(n, 24), c->F), dim3(256), trsm_smem_bytes(k, 24, 3), c->stream, v));
            else if (trsm_smem_bytes(k, 24, 2) <= smemMax)
                CK(launch_pdl(k_trsm_slab<24, 2>, dim3(cdiv(n, 24), c->F), dim3(256), trsm_smem_bytes(k, 24, 2), c->stream, v));
            else if (trsm_smem_bytes(k, 16, 4) <= smemMax && maxStages >= 4)
                CK(launch_pdl(k_trsm_slab<16, 4>, dim3(cdiv(n, 16), c->F), dim3(256), trsm_smem_bytes(k, 16, 4), c->stream, v));
            else
                CK(launch_pdl(k_trsm_slab<16, 2>, dim3(cdiv(n, 16), c->F), dim3(256), smem16, c->stream, v));
            count_launch(c);
            if (c->schain_eff >= 3) {   // next generation of the chain's flags (after everything that follows the chain)
                k_chain_finish<<<cdiv(c->F, 128), 128, 0, c->stream>>>(c->chainCtl, c->nbMax, c->F);
                count_launch(c);
            }
        } else if (c->F == 1 && c->tmaEncode && c->force_generic != 1) {
            // large k (the slab of the TRSM does not fit in shared memory): S-chain, one launch per block step, then the
            // blocked left-looking TRSM on the global-memory resident B (ekf_gemm_tma.cuh): per 64-row block one long
            // contraction over the rows already solved and one 64-row product with the inverse of the diagonal block
            c->schain_eff = 0;
            { int rcC = launch_schain(c, k); if (rcC != EKFB_OK) return rcC; }
            const int kpad = steps * kNB;
            if (kpad > k)   // padding rows of the last block: finite (they meet the identity padding of Uinv)
                CK(cudaMemsetAsync(v.Bu + (size_t)k * c->ld, 0, sizeof(double) * (size_t)(kpad - k) * c->ld, c->stream));
            typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            auto encode = [&](CUtensorMap* m, void* base, cuuint64_t cols, cuuint64_t rows, cuuint64_t pitchDoubles, cuuint32_t boxRows) {
                const cuuint64_t dims[2] = {cols, rows};
                const cuuint64_t strides[1] = {pitchDoubles * sizeof(double)};
                const cuuint32_t box[2] = {16u, boxRows}, es[2] = {1, 1};
                return ((EncodeFn)c->tmaEncode)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            };
            GemmMaps mU, mI;   // contraction with rows of U (factor buffer) / with the inverse of a diagonal block
            bool ok = encode(&mU.A, v.Sf, (cuuint64_t)c->ldS, (cuuint64_t)c->kmax, (cuuint64_t)c->ldS, 16) &&
                      encode(&mU.B, v.Bu, (cuuint64_t)n, (cuuint64_t)c->kmax, (cuuint64_t)c->ld, 16) &&
                      encode(&mU.C, v.Bu, (cuuint64_t)n, (cuuint64_t)c->kmax, (cuuint64_t)c->ld, 64) &&
                      encode(&mI.A, v.Uinv, (cuuint64_t)kNB, (cuuint64_t)c->kmax, (cuuint64_t)kNB, 16);
            if (!ok) {
                g_err = "cuTensorMapEncodeTiled failed for the global-memory TRSM operands";
                return EKFB_ERR_CUDA;
            }
            mI.B = mU.B;
            mI.C = mU.C;
            const dim3 tgrid(cdiv(n, 64));
            for (int J = 0; J < steps; ++J) {
                const int J0 = J * kNB;
                if (J > 0) CK(launch_k(c, k_gemm_tn_tma, tgrid, dim3(160), (size_t)kGtSmemBytes, mU, J0, 0, 0, J0, J0, 1));
                CK(launch_k(c, k_gemm_tn_tma, tgrid, dim3(160), (size_t)kGtSmemBytes, mI, 0, J0, J0, J0, kNB, 0));
                count_launch(c, J > 0 ? 2 : 1);
            }
            CK(launch_k(c, k_wy, dim3(cdiv(n, 256)), dim3(256), 0, v));   // dx = W y; clears the rows the downdate reads beyond k
            CK(launch_k(c, k_state_apply, dim3(cdiv(n, 256), c->F), dim3(256), 0, v));
            count_launch(c, 2);
        } else {
            // generic path (very large k): right-looking over the whole augmented matrix
            usedGeneric = true;
            for (int J = 0; J < steps; ++J) {
                const int J1 = (J + 1) * kNB;
                const int nS = k > J1 ? cdiv(k - J1, kPanelCols) : 0;
                const int nB = cdiv(n + 1, kPanelCols);
                k_chol_panel<<<dim3(nS + nB, c->F), 128, 0, c->stream>>>(v, J);
                count_launch(c);
                if (k > J1) {
                    const int mt = cdiv(k - J1, kTM);
                    k_gemm_tn<0><<<dim3(cdiv(k - J1, kTN) + cdiv(n + 1, kTN), mt, c->F), 256, kGemmSmemBytes, c->stream>>>(v, J);
                    count_launch(c);
                }
            }
        }
        if (usedGeneric) {   // (every other path applies the state correction in the last block of its TRSM kernel)
            CK(launch_k(c, k_state_apply, dim3(cdiv(n, 256), c->F), dim3(256), 0, v));   // + quaternion normalisation (block 0)
            count_launch(c);
        }
    }
    {
        GroupScope gs(c, G_DOWNDATE);
        int rcD = launch_downdate(c, n, !usedGeneric);
        if (rcD != EKFB_OK) return rcD;
        CK(launch_k(c, k_quat_cov, dim3(cdiv(n, 256), c->F), dim3(256), 0, v));
        count_launch(c, 2);
    }
    CK(cudaGetLastError());
    return EKFB_OK;
}

extern "C" int ekfb_update(ekfb_handle c, int which)
{
    REQUIRE(c, "null handle");
    REQUIRE(which == 0 || which == 1, "which must be 0 (low innovation) or 1 (high innovation)");
    CK(cudaSetDevice(c->device));
    return run_update(c, which);
}

static int rescue_launch(ekfb_ctx* c, int* seq)
{
    // the re-prediction of the outliers; the block that finishes last applies the chi-square gate and publishes the counters
    *seq = ++c->dims_seq;
    return launch_measure(c, 1, *seq, c->in_step ? 2 : 0);   // inside ekfb_step: + the map-feature bookkeeping of the frame
}

extern "C" int ekfb_rescue(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    GroupScope gs(c, G_RESCUE);
    int seq = 0;
    int rc = rescue_launch(c, &seq);
    if (rc != EKFB_OK) return rc;
    return wait_published_dims(c, seq);
}

extern "C" int ekfb_update_map_features(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    GroupScope gs(c, G_MISC);
    const int N = max_of(c->hN);
    if (N == 0) return EKFB_OK;
    CK(launch_k(c, k_update_map_features, dim3(cdiv(N, 256), c->F), dim3(256), 0, c->v));
    count_launch(c);
    CK(cudaGetLastError());
    return EKFB_OK;
}

static void fill_info(const int* d, ekfb_frame_info* o);

// ---- map management (kernels in ekf_map.cuh) ----------------------------------------------
static int ensure_map_buffers(ekfb_ctx* c)
{
    if (c->map_ready) return EKFB_OK;
    DevView& v = c->v;
    const size_t F = c->F, N = c->Nmax;
    ALLOC(v.P2, F * c->nmax * c->ld); ALLOC(v.x2, F * c->ld);
    ALLOC(v.ftype2, F * N); ALLOC(v.foff2, F * N); ALLOC(v.desc2, F * N * 32); ALLOC(v.tpred2, F * N); ALLOC(v.tmatch2, F * N);
    ALLOC(v.rowsrc, F * c->nmax); ALLOC(v.mapflag, F * N); ALLOC(v.convJ, F * 18);
    ALLOC(v.addJ, N * kAddJ); ALLOC(v.adduv, N * 2); ALLOC(v.adddesc, N * 32);
    ALLOC(c->mask2, F * (size_t)v.W * v.H);
    c->map_ready = true;
    return EKFB_OK;
}

extern "C" int ekfb_map_management(ekfb_handle c, const ekfb_map_policy* pol, ekfb_map_result* out)
{
    REQUIRE(c && pol, "null argument");
    CK(cudaSetDevice(c->device));
    int rc = ensure_map_buffers(c);
    if (rc != EKFB_OK) return rc;
    GroupScope gs(c, G_MISC);
    DevView& v = c->v;
    MapPolicy mp;
    mp.min_matches = pol->min_matches_per_image; mp.max_features = pol->max_map_features_count;
    mp.max_size = pol->max_map_size; mp.always_remove_unseen = pol->always_remove_unseen;
    mp.good_pct = pol->good_feature_matching_percent; mp.linearity_thr = pol->linearity_index_threshold;
    k_map_plan<<<c->F, 256, 0, c->stream>>>(v, mp);
    count_launch(c);
    CK(cudaGetLastError());
    if ((rc = read_dims(c)) != EKFB_OK) return rc;
    bool changed = false, needNew = false;
    int nnMax = 13;
    for (int f = 0; f < c->F; ++f) {
        const int* d = c->h_dims + (size_t)f * D_STRIDE;
        changed = changed || d[D_MAP_CHANGED];
        needNew = needNew || d[D_MAP_NEEDED] > 0;
        nnMax = std::max(nnMax, d[D_MAP_NEW_N]);
    }
    c->mask2_valid = false;
    if (needNew && max_of(c->hN) > 0) {
        // buildImageMask (E/DetectNewImageFeatures.cpp:101-122): white image, every prediction's ellipse in black.  Built
        // here because the predictions are indexed by the feature numbering of this frame's measurement.
        CK(cudaMemsetAsync(c->mask2, 255, (size_t)c->F * v.W * v.H, c->stream));
        const int rasterSmem = 4 * (int)(sizeof(RasterScratch) + sizeof(int) * 2 * (size_t)v.H);
        CK(launch_k(c, k_mask_raster, dim3(cdiv(max_of(c->hN), 4), c->F), dim3(128), rasterSmem, v, c->mask2, 2 * (v.W + v.H), 0));
        count_launch(c);
        c->mask2_valid = true;
    } else if (needNew) {
        CK(cudaMemsetAsync(c->mask2, 255, (size_t)c->F * v.W * v.H, c->stream));
        c->mask2_valid = true;
    }
    if (changed) {
        k_map_gather<<<dim3(cdiv(nnMax, 32), cdiv(nnMax, 32), c->F), dim3(32, 8), 0, c->stream>>>(v);
        k_map_convert<<<dim3(cdiv(nnMax, 256), c->F), 256, 0, c->stream>>>(v);
        k_map_commit<<<cdiv(c->F, 128), 128, 0, c->stream>>>(v);
        count_launch(c, 3);
        if (c->ncc_ready) {   // NCC templates and anchors follow their features
            const int Nold = max_of(c->hN);
            k_ncc_compact<<<c->F, 256, sizeof(int) * std::max(Nold, 1), c->stream>>>(v, c->ncc_tmpl, c->ncc_tmpl2, c->ncc_anchor, c->ncc_anchor2, Nold);
            count_launch(c);
            std::swap(c->ncc_tmpl, c->ncc_tmpl2);
            std::swap(c->ncc_anchor, c->ncc_anchor2);
        }
        CK(cudaGetLastError());
        std::swap(v.P, v.P2); std::swap(v.x, v.x2); std::swap(v.ftype, v.ftype2); std::swap(v.foff, v.foff2);
        std::swap(v.desc, v.desc2); std::swap(v.tpred, v.tpred2); std::swap(v.tmatch, v.tmatch2);
    }
    for (int f = 0; f < c->F; ++f) {
        int* d = c->h_dims + (size_t)f * D_STRIDE;
        if (changed) {
            c->hn[f] = d[D_N_STATE] = d[D_MAP_NEW_N];
            c->hN[f] = d[D_N_FEAT] = d[D_MAP_NEW_NF];
        }
        if (out) {
            out[f].n = c->hn[f]; out[f].n_features = c->hN[f]; out[f].n_removed_bad = d[D_MAP_NBAD];
            out[f].n_removed_unseen = d[D_MAP_NUNSEEN]; out[f].converted = d[D_MAP_CONVERT];
            out[f].new_features_needed = d[D_MAP_NEEDED];
        }
    }
    return EKFB_OK;
}

extern "C" int ekfb_get_removed_flags(ekfb_handle c, int f, int n_features_before, uint8_t* flags)
{
    REQUIRE(c && flags, "null argument");
    REQUIRE(f >= 0 && f < c->F && n_features_before >= 0 && n_features_before <= c->Nmax, "bad filter index or count");
    REQUIRE(c->map_ready, "ekfb_map_management has not been called");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(flags, c->v.mapflag + (size_t)f * c->Nmax, n_features_before, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_add_features(ekfb_handle c, int f, int count, const double* uv, const uint8_t* desc)
{
    REQUIRE(c && (count == 0 || (uv && desc)), "null argument");
    REQUIRE(f >= 0 && f < c->F && count >= 0, "bad filter index or count");
    if (count == 0) return EKFB_OK;
    const int n0 = c->hn[f], N0 = c->hN[f];
    if (N0 + count > c->Nmax || n0 + 6 * count > c->nmax) {
        g_err = "new features exceed the capacity reserved by ekfb_create";
        return EKFB_ERR_CAPACITY;
    }
    CK(cudaSetDevice(c->device));
    int rc = ensure_map_buffers(c);
    if (rc != EKFB_OK) return rc;
    GroupScope gs(c, G_MISC);
    DevView& v = c->v;
    CK(cudaMemcpyAsync(v.adduv, uv, sizeof(double) * 2 * count, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(v.adddesc, desc, (size_t)32 * count, cudaMemcpyHostToDevice, c->stream));
    k_add_prepare<<<count, 32, 0, c->stream>>>(v, f, n0, N0, count, c->prm.pixel_error_x, c->prm.pixel_error_y,
                                                c->prm.init_inv_depth_rho, c->prm.inverse_depth_rho_sd);
    k_add_cov<<<dim3(cdiv(n0 + 6 * count, 256), count), 256, 0, c->stream>>>(v, f, n0, count);
    count_launch(c, 2);
    if (c->ncc_ready && c->ncc_has_image[f]) {
        // NCC appearance of the new features: templates cut from the frame's pyramid at their pixels, and the camera pose
        // they were seen from (ekf_ncc.cuh)
        NccView nv = c->ncc;
        for (int l = 0; l < kNccLevels; ++l) nv.img[l] = c->ncc_img[l] + (size_t)f * c->ncc_level_bytes[l];
        k_ncc_capture<<<count, 128, 0, c->stream>>>(v, nv, f, N0, count, v.adduv, c->ncc_tmpl + (size_t)f * c->Nmax * kNccLevels * 128,
                                                    c->ncc_anchor + (size_t)f * c->Nmax * kNccAnchor);
        count_launch(c);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));   // uv / desc may be pageable
    c->hn[f] = n0 + 6 * count;
    c->hN[f] = N0 + count;
    c->h_dims[(size_t)f * D_STRIDE + D_N_STATE] = c->hn[f];
    c->h_dims[(size_t)f * D_STRIDE + D_N_FEAT] = c->hN[f];
    return EKFB_OK;
}

extern "C" int ekfb_get_feature_layout(ekfb_handle c, int f, int32_t* type, int32_t* off)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const int N = c->hN[f];
    if (type) CK(cudaMemcpyAsync(type, c->v.ftype + (size_t)f * c->Nmax, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
    if (off) CK(cudaMemcpyAsync(off, c->v.foff + (size_t)f * c->Nmax, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_set_hit_counters(ekfb_handle c, int f, const int32_t* tp, const int32_t* tm)
{
    REQUIRE(c && tp && tm, "null argument");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const int N = c->hN[f];
    CK(cudaMemcpyAsync(c->v.tpred + (size_t)f * c->Nmax, tp, sizeof(int) * N, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v.tmatch + (size_t)f * c->Nmax, tm, sizeof(int) * N, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

// what the host's new-feature selection reads, in one round trip: the predictions of this frame's measurement (indexed by
// the feature numbering BEFORE ekfb_map_management, which leaves these arrays alone) and the new-feature mask
extern "C" int ekfb_get_new_feature_inputs(ekfb_handle c, int f, int n_features_before, uint8_t* predicted, double* hpred, uint8_t* mask)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F && n_features_before >= 0 && n_features_before <= c->Nmax, "bad filter index or count");
    REQUIRE(!mask || (c->map_ready && c->mask2_valid), "no new-feature mask: the last ekfb_map_management asked for no new features");
    CK(cudaSetDevice(c->device));
    const size_t fo = (size_t)f * c->Nmax, wh = (size_t)c->v.W * c->v.H;
    if (predicted && n_features_before)
        CK(cudaMemcpyAsync(predicted, c->v.vis + fo, n_features_before, cudaMemcpyDeviceToHost, c->stream));
    if (hpred && n_features_before)
        CK(cudaMemcpyAsync(hpred, c->v.h + fo * 2, sizeof(double) * 2 * n_features_before, cudaMemcpyDeviceToHost, c->stream));
    if (mask) CK(cudaMemcpyAsync(mask, c->mask2 + (size_t)f * wh, wh, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

// the host mirror of one filter in one round trip: state vector, feature layout, descriptors, hit counters, result record
extern "C" int ekfb_get_map_snapshot(ekfb_handle c, int f, double* x, int32_t* type, int32_t* off, uint8_t* desc, int32_t* tp,
                                     int32_t* tm, ekfb_record* rec)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const int n = c->hn[f], N = c->hN[f];
    const size_t fo = (size_t)f * c->Nmax;
    if (rec) {
        k_write_records<<<c->F, 192, 0, c->stream>>>(c->v, c->d_rec);
        count_launch(c);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c->h_rec + f, c->d_rec + f, sizeof(RecordDev), cudaMemcpyDeviceToHost, c->stream));
    }
    if (x) CK(cudaMemcpyAsync(x, c->v.x + (size_t)f * c->ld, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (N > 0) {
        if (type) CK(cudaMemcpyAsync(type, c->v.ftype + fo, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
        if (off) CK(cudaMemcpyAsync(off, c->v.foff + fo, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
        if (desc) CK(cudaMemcpyAsync(desc, c->v.desc + fo * 32, (size_t)N * 32, cudaMemcpyDeviceToHost, c->stream));
        if (tp) CK(cudaMemcpyAsync(tp, c->v.tpred + fo, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
        if (tm) CK(cudaMemcpyAsync(tm, c->v.tmatch + fo, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    if (rec) std::memcpy(rec, c->h_rec + f, sizeof(ekfb_record));
    return EKFB_OK;
}

// the counters of the last frame from the host mirror, without a device round trip: valid right after a call that
// synchronised them (ekfb_step / ekfb_rescue / ekfb_map_management / ekfb_get_frame_info); `status` may lag behind the
// high-innovation update (ekfb_get_frame_info reads it from the device)
extern "C" int ekfb_peek_frame_info(ekfb_handle c, int f, ekfb_frame_info* info)
{
    REQUIRE(c && info, "null argument");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    fill_info(c->h_dims + (size_t)f * D_STRIDE, info);
    return EKFB_OK;
}

extern "C" int ekfb_get_new_feature_mask(ekfb_handle c, int f, uint8_t* mask)
{
    REQUIRE(c && mask, "null argument");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    REQUIRE(c->map_ready && c->mask2_valid, "no new-feature mask: the last ekfb_map_management asked for no new features");
    CK(cudaSetDevice(c->device));
    const size_t wh = (size_t)c->v.W * c->v.H;
    CK(cudaMemcpyAsync(mask, c->mask2 + (size_t)f * wh, wh, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_raster_ellipse(ekfb_handle c, int W, int H, double cx, double cy, const double* S, int max_axes, int value,
                                   uint8_t* img)
{
    REQUIRE(c && S && img && W > 0 && H > 0, "bad argument");
    CK(cudaSetDevice(c->device));
    uint8_t* d = nullptr;
    CK(cudaMalloc(&d, (size_t)W * H));
    const int smem = (int)(sizeof(RasterScratch) + sizeof(int) * 2 * (size_t)H);
    cudaError_t e = cudaMemcpyAsync(d, img, (size_t)W * H, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_raster_one, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) {
        k_raster_one<<<1, 32, smem, c->stream>>>(d, W, H, cx, cy, S[0], S[1], S[2], S[3], max_axes, value);
        count_launch(c);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(img, d, (size_t)W * H, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    CK(e);
    return EKFB_OK;
}

// ---- NCC active search (north-star path, ekf_ncc.cuh) ---------------------------------------------------------
static int ensure_ncc(ekfb_ctx* c)
{
    if (c->ncc_ready) return EKFB_OK;
    DevView& v = c->v;
    NccView& nv = c->ncc;
    int W = v.W, H = v.H;
    for (int l = 0; l < kNccLevels; ++l) {
        nv.W[l] = W; nv.H[l] = H; nv.pitch[l] = rup(W, 16);
        c->ncc_level_bytes[l] = (size_t)nv.pitch[l] * H;
        ALLOC(c->ncc_img[l], (size_t)c->F * c->ncc_level_bytes[l]);
        W /= 2; H /= 2;
    }
    ALLOC(c->ncc_tmpl, (size_t)c->F * c->Nmax * kNccLevels * 128);
    ALLOC(c->ncc_tmpl2, (size_t)c->F * c->Nmax * kNccLevels * 128);
    ALLOC(c->ncc_anchor, (size_t)c->F * c->Nmax * kNccAnchor);
    ALLOC(c->ncc_anchor2, (size_t)c->F * c->Nmax * kNccAnchor);
    c->ncc_has_image.assign(c->F, 0);
    // tensor maps of the pyramid levels (one window load per level and feature instead of 36 bulk row copies); a level smaller
    // than the box, or a failed encode, keeps the bulk-copy form
    c->ncc_maps.assign(c->F, NccMaps());
    for (int l = 0; l < kNccLevels; ++l) {
        bool ok = c->tmaEncode != nullptr && nv.W[l] >= kNccBoxW && nv.H[l] >= kNccBoxH;
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        for (int f = 0; ok && f < c->F; ++f) {
            const cuuint64_t dims[2] = {(cuuint64_t)nv.W[l], (cuuint64_t)nv.H[l]};
            const cuuint64_t strides[1] = {(cuuint64_t)nv.pitch[l]};
            const cuuint32_t box[2] = {(cuuint32_t)kNccBoxW, (cuuint32_t)kNccBoxH}, es[2] = {1, 1};
            ok = ((EncodeFn)c->tmaEncode)(&c->ncc_maps[f].m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->ncc_img[l] + (size_t)f * c->ncc_level_bytes[l],
                                          dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        c->ncc_tma_level[l] = ok ? 1 : 0;
    }
    ALLOC(nv.score, (size_t)c->F * c->Nmax);
    ALLOC(nv.level, (size_t)c->F * c->Nmax);
    nv.ncc_min = 0.8;
    ALLOC(c->fe_score, (size_t)c->F * c->ncc_level_bytes[0]);
    ALLOC(c->fe_rows, (size_t)c->F * 2 * v.H);
    ALLOC(c->fe_color, (size_t)v.W * v.H * 4);
    ALLOC(c->fe_count, (size_t)c->F);
    CK(cudaMemcpyToSymbolAsync(c_brief, kBriefPattern, sizeof(kBriefPattern), 0, cudaMemcpyHostToDevice, c->stream));
    c->ncc_ready = true;
    return EKFB_OK;
}

extern "C" int ekfb_ncc_set_image(ekfb_handle c, int f, const uint8_t* gray, int stride)
{
    REQUIRE(c && gray, "null argument");
    REQUIRE(f >= 0 && f < c->F && stride >= c->v.W, "bad filter index or stride");
    CK(cudaSetDevice(c->device));
    int rc = ensure_ncc(c);
    if (rc != EKFB_OK) return rc;
    NccView& nv = c->ncc;
    uint8_t* L[kNccLevels];
    for (int l = 0; l < kNccLevels; ++l) L[l] = c->ncc_img[l] + (size_t)f * c->ncc_level_bytes[l];
    CK(cudaMemcpy2DAsync(L[0], nv.pitch[0], gray, stride, nv.W[0], nv.H[0], cudaMemcpyHostToDevice, c->stream));
    for (int l = 1; l < kNccLevels; ++l) {
        k_pyr_down<<<dim3(cdiv(nv.W[l], 32), cdiv(nv.H[l], 8)), dim3(32, 8), 0, c->stream>>>(L[l - 1], nv.pitch[l - 1], L[l], nv.pitch[l],
                                                                                          nv.W[l], nv.H[l]);
        count_launch(c);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));   // the caller's image may be pageable
    c->ncc_has_image[f] = 1;
    return EKFB_OK;
}

extern "C" int ekfb_set_image(ekfb_handle c, int f, const uint8_t* gray, int stride) { return ekfb_ncc_set_image(c, f, gray, stride); }

// the frame as the callers of the reference hold it: 8-bit interleaved BGR (desktop, FileSequenceImageGenerator.cpp:82) or BGRA
// (Android, EKFNative.cpp:134-137); converted to grey on the device, then as ekfb_set_image
extern "C" int ekfb_set_image_color(ekfb_handle c, int f, const uint8_t* pixels, int stride, int channels)
{
    REQUIRE(c && pixels, "null argument");
    REQUIRE(channels == 1 || channels == 3 || channels == 4, "channels must be 1, 3 or 4");
    if (channels == 1) return ekfb_ncc_set_image(c, f, pixels, stride);
    REQUIRE(f >= 0 && f < c->F && stride >= c->v.W * channels, "bad filter index or stride");
    CK(cudaSetDevice(c->device));
    int rc = ensure_ncc(c);
    if (rc != EKFB_OK) return rc;
    NccView& nv = c->ncc;
    const int W = nv.W[0], H = nv.H[0];
    uint8_t* L[kNccLevels];
    for (int l = 0; l < kNccLevels; ++l) L[l] = c->ncc_img[l] + (size_t)f * c->ncc_level_bytes[l];
    CK(cudaMemcpy2DAsync(c->fe_color, (size_t)W * channels, pixels, stride, (size_t)W * channels, H, cudaMemcpyHostToDevice, c->stream));
    k_bgr_to_gray<<<dim3(cdiv(W, 32), cdiv(H, 8)), dim3(32, 8), 0, c->stream>>>(c->fe_color, W * channels, channels, L[0], nv.pitch[0], W, H);
    count_launch(c);
    for (int l = 1; l < kNccLevels; ++l) {
        k_pyr_down<<<dim3(cdiv(nv.W[l], 32), cdiv(nv.H[l], 8)), dim3(32, 8), 0, c->stream>>>(L[l - 1], nv.pitch[l - 1], L[l], nv.pitch[l],
                                                                                          nv.W[l], nv.H[l]);
        count_launch(c);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));   // the caller's frame may be pageable
    return EKFB_OK;
}

// detector + descriptor on the device for the image of ekfb_set_image: the keypoints become the frame's front-end output
extern "C" int ekfb_detect_keypoints(ekfb_handle c, int f, int threshold, int32_t* n_kp)
{
    REQUIRE(c && c->ncc_ready, "ekfb_set_image first");
    REQUIRE(f >= 0 && f < c->F && threshold >= 1 && threshold <= 254, "bad filter index or threshold");
    CK(cudaSetDevice(c->device));
    GroupScope gs(c, G_MATCH);
    NccView& nv = c->ncc;
    const int W = nv.W[0], H = nv.H[0], pitch = nv.pitch[0];
    const uint8_t* img = c->ncc_img[0] + (size_t)f * c->ncc_level_bytes[0];
    uint8_t* score = c->fe_score + (size_t)f * c->ncc_level_bytes[0];
    int* rowCount = c->fe_rows + (size_t)f * 2 * H;
    int* rowOff = rowCount + H;
    int* count = c->fe_count + f;
    float* dxy = c->d_kpxy + (size_t)f * c->Kpmax * 2;
    uint8_t* dds = c->d_kpdesc + (size_t)f * c->Kpmax * 32;
    k_fast_score<<<dim3(cdiv(W, 32), cdiv(H, 8)), dim3(32, 8), 0, c->stream>>>(img, pitch, W, H, threshold, score);
    k_fast_rows<<<cdiv(H, 8), 256, 0, c->stream>>>(score, pitch, W, H, rowCount, rowOff, dxy, c->Kpmax, 0);
    k_fast_scan<<<1, 1024, 0, c->stream>>>(rowCount, rowOff, H, c->Kpmax, count);
    k_fast_rows<<<cdiv(H, 8), 256, 0, c->stream>>>(score, pitch, W, H, rowCount, rowOff, dxy, c->Kpmax, 1);
    k_brief<<<cdiv(c->Kpmax, 8), 256, 0, c->stream>>>(img, pitch, W, H, dxy, count, dds);
    count_launch(c, 5);
    CK(cudaGetLastError());
    int hcount = 0;
    CK(cudaMemcpyAsync(&hcount, count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->hKp[f] = hcount;
    c->h_kpxy_ptr[f] = dxy;
    c->h_kpdesc_ptr[f] = dds;
    c->h_dims[(size_t)f * D_STRIDE + D_N_KP] = hcount;
    CK(cudaMemcpyAsync(c->d_kpxy_ptr + f, c->h_kpxy_ptr + f, sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_kpdesc_ptr + f, c->h_kpdesc_ptr + f, sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v.dims + (size_t)f * D_STRIDE + D_N_KP, c->h_dims + (size_t)f * D_STRIDE + D_N_KP, sizeof(int),
                       cudaMemcpyHostToDevice, c->stream));
    if (n_kp) *n_kp = hcount;
    return EKFB_OK;
}

extern "C" int ekfb_get_keypoints(ekfb_handle c, int f, float* xy, uint8_t* desc)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const int n = c->hKp[f];
    if (xy && n) CK(cudaMemcpyAsync(xy, c->h_kpxy_ptr[f], sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    if (desc && n) CK(cudaMemcpyAsync(desc, c->h_kpdesc_ptr[f], (size_t)32 * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_ncc_get_level(ekfb_handle c, int f, int level, uint8_t* out, int32_t* w, int32_t* h)
{
    REQUIRE(c && c->ncc_ready, "no image set");
    REQUIRE(f >= 0 && f < c->F && level >= 0 && level < kNccLevels, "bad filter index or level");
    CK(cudaSetDevice(c->device));
    NccView& nv = c->ncc;
    if (w) *w = nv.W[level];
    if (h) *h = nv.H[level];
    if (out)
        CK(cudaMemcpy2DAsync(out, nv.W[level], c->ncc_img[level] + (size_t)f * c->ncc_level_bytes[level], nv.pitch[level], nv.W[level],
                             nv.H[level], cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_ncc_set_templates(ekfb_handle c, int f, int first_feature, int count, const uint8_t* templates)
{
    REQUIRE(c && templates, "null argument");
    REQUIRE(f >= 0 && f < c->F && first_feature >= 0 && count >= 0 && first_feature + count <= c->Nmax, "bad filter index or range");
    CK(cudaSetDevice(c->device));
    int rc = ensure_ncc(c);
    if (rc != EKFB_OK) return rc;
    uint8_t* dst = c->ncc_tmpl + ((size_t)f * c->Nmax + first_feature) * kNccLevels * 128;
    CK(cudaMemcpy2DAsync(dst, 128, templates, kNccPP, kNccPP, (size_t)count * kNccLevels, cudaMemcpyHostToDevice, c->stream));
    // caller-supplied templates carry no anchor: they are compared as they are (no warp)
    CK(cudaMemsetAsync(c->ncc_anchor + ((size_t)f * c->Nmax + first_feature) * kNccAnchor, 0, sizeof(double) * kNccAnchor * count, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_ncc_get_templates(ekfb_handle c, int f, int first_feature, int count, uint8_t* templates, double* anchors)
{
    REQUIRE(c && c->ncc_ready, "no NCC data");
    REQUIRE(f >= 0 && f < c->F && first_feature >= 0 && count >= 0 && first_feature + count <= c->Nmax, "bad filter index or range");
    CK(cudaSetDevice(c->device));
    if (templates)
        CK(cudaMemcpy2DAsync(templates, kNccPP, c->ncc_tmpl + ((size_t)f * c->Nmax + first_feature) * kNccLevels * 128, 128, kNccPP,
                             (size_t)count * kNccLevels, cudaMemcpyDeviceToHost, c->stream));
    if (anchors)
        CK(cudaMemcpyAsync(anchors, c->ncc_anchor + ((size_t)f * c->Nmax + first_feature) * kNccAnchor, sizeof(double) * kNccAnchor * count,
                           cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_ncc_set_anchors(ekfb_handle c, int f, int first_feature, int count, const double* anchors)
{
    REQUIRE(c && anchors, "null argument");
    REQUIRE(f >= 0 && f < c->F && first_feature >= 0 && count >= 0 && first_feature + count <= c->Nmax, "bad filter index or range");
    CK(cudaSetDevice(c->device));
    int rc = ensure_ncc(c);
    if (rc != EKFB_OK) return rc;
    CK(cudaMemcpyAsync(c->ncc_anchor + ((size_t)f * c->Nmax + first_feature) * kNccAnchor, anchors, sizeof(double) * kNccAnchor * count,
                       cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_ncc_set_threshold(ekfb_handle c, double ncc_min)
{
    REQUIRE(c, "null handle");
    c->ncc_min_score = ncc_min;
    return EKFB_OK;
}

extern "C" int ekfb_match_ncc(ekfb_handle c, double ncc_min)
{
    REQUIRE(c && c->ncc_ready, "ekfb_ncc_set_image / ekfb_ncc_set_templates first");
    CK(cudaSetDevice(c->device));
    GroupScope gs(c, G_MATCH);
    for (int f = 0; f < c->F; ++f) {
        if (c->hN[f] == 0) continue;
        NccView nv = c->ncc;
        nv.tmpl = c->ncc_tmpl + (size_t)f * c->Nmax * kNccLevels * 128;
        nv.score = c->ncc.score + (size_t)f * c->Nmax;
        nv.level = c->ncc.level + (size_t)f * c->Nmax;
        nv.ncc_min = ncc_min;
        nv.anchor = c->ncc_anchor + (size_t)f * c->Nmax * kNccAnchor;
        nv.warp = c->ncc_warp;
        for (int l = 0; l < kNccLevels; ++l) nv.img[l] = c->ncc_img[l] + (size_t)f * c->ncc_level_bytes[l];
        for (int l = 0; l < kNccLevels; ++l) nv.tmaLevel[l] = c->ncc_tma_window ? c->ncc_tma_level[l] : 0;
        k_search_ncc<<<c->hN[f], 128, 0, c->stream>>>(c->v, nv, c->ncc_maps[f], f);
        count_launch(c);
    }
    CK(launch_k(c, k_after_match, dim3(c->F), dim3(256), 0, c->v));
    count_launch(c);
    CK(cudaGetLastError());
    return EKFB_OK;
}

extern "C" int ekfb_ncc_get_scores(ekfb_handle c, int f, double* score, int32_t* level)
{
    REQUIRE(c && c->ncc_ready, "no NCC search has run");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const int N = c->hN[f];
    if (score) CK(cudaMemcpyAsync(score, c->ncc.score + (size_t)f * c->Nmax, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    if (level) CK(cudaMemcpyAsync(level, c->ncc.level + (size_t)f * c->Nmax, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

// the view of filters f0 .. f0 + cnt - 1: every filter-major array starts at its f0-th slice
static DevView lane_view(const ekfb_ctx* c, int f0, int cnt)
{
    DevView s = c->v;
    s.F = cnt;
    const size_t f = (size_t)f0, N = (size_t)c->Nmax;
#define LANE_OFF(p, stride) do { if (s.p) s.p += f * (size_t)(stride); } while (0)
    LANE_OFF(x, c->ld); LANE_OFF(P, (size_t)c->nmax * c->ld);
    LANE_OFF(ftype, N); LANE_OFF(foff, N); LANE_OFF(desc, N * 32); LANE_OFF(tpred, N); LANE_OFF(tmatch, N); LANE_OFF(dims, D_STRIDE);
    LANE_OFF(vis, N); LANE_OFF(h, N * 2); LANE_OFF(Si, N * 4); LANE_OFF(Hx, N * 14); LANE_OFF(Hf, N * 12); LANE_OFF(ellax, N * 2); LANE_OFF(ellang, N);
    LANE_OFF(vis2, N); LANE_OFF(h2, N * 2); LANE_OFF(Si2, N * 4); LANE_OFF(Hx2, N * 14); LANE_OFF(Hf2, N * 12);
    LANE_OFF(mflag, N); LANE_OFF(z, N * 2); LANE_OFF(mkp, N); LANE_OFF(mdist, N); LANE_OFF(mlist, N);
    LANE_OFF(inl, N); LANE_OFF(outl, N); LANE_OFF(resc, N); LANE_OFF(ulist, N);
    LANE_OFF(kpxy, 1); LANE_OFF(kpdesc, 1); LANE_OFF(kpok, c->Kpmax); LANE_OFF(mask, (size_t)s.W * s.H);
    LANE_OFF(hypcount, N); LANE_OFF(hypsup, N * c->supWords);
    LANE_OFF(Bu, (size_t)c->kmax * c->ld); LANE_OFF(S, (size_t)c->kmax * c->ldS); LANE_OFF(Sf, (size_t)c->kmax * c->ldS); LANE_OFF(dx, c->ld);
    LANE_OFF(Jq, 16); LANE_OFF(Uinv, (size_t)(c->kmax / kNB) * kNB * kNB);
    LANE_OFF(hostDims, D_STRIDE); LANE_OFF(hostFlag, 1);
#undef LANE_OFF
    return s;
}

// while alive, the handle IS the lane: view, filter count, stream and the host mirrors are the lane's, so the phase functions
// run unchanged on the lane's filters
struct LaneScope {
    ekfb_ctx* c;
    Lane* L;
    DevView v0;
    int F0;
    cudaStream_t s0;
    LaneScope(ekfb_ctx* c_, Lane* L_) : c(c_), L(L_), v0(c_->v), F0(c_->F), s0(c_->stream)
    {
        c->v = L->v; c->F = L->F; c->stream = L->stream;
        c->ddQueueNext = L->idx;
        std::swap(c->hn, L->hn); std::swap(c->hN, L->hN); std::swap(c->hKp, L->hKp);
        shift(+1);
    }
    ~LaneScope()
    {
        shift(-1);
        std::swap(c->hn, L->hn); std::swap(c->hN, L->hN); std::swap(c->hKp, L->hKp);
        c->v = v0; c->F = F0; c->stream = s0;
        c->ddQueueNext = 0;
    }
    void shift(int sign)
    {
        const ptrdiff_t f = (ptrdiff_t)sign * L->f0;
        c->h_dims += f * D_STRIDE; c->h_dims_zc += f * D_STRIDE; c->h_flag_zc += f;
        c->chainCtl += f * (ptrdiff_t)chain_ctl_ints(c->nbMax);
    }
};

static int lanes_wanted(const ekfb_ctx* c)
{
    if (c->prof) return 1;   // the per-group timers bracket one stream
    if (c->matcher == 1) return 1;   // the NCC search indexes its own per-filter arrays (not part of the lane views)
    const int want = c->lanes_opt > 0 ? c->lanes_opt : (c->F >= 8 ? 2 : 1);
    return std::max(1, std::min(std::min(want, 8), c->F));
}

// The frame of a batched handle as `nl` lanes on `nl` streams, interleaved phase by phase by this one host thread: while the host
// waits for lane A's counters (RANSAC, rescue) lane B's kernels are already queued, and on the device the short latency-bound
// kernels of one lane run beside the covariance downdate of the other.  Same kernels, same per-filter results.
static int step_lanes(ekfb_ctx* c, int nl)
{
    CK(cudaSetDevice(c->device));
    if ((int)c->lanes.size() != nl) {
        for (Lane& L : c->lanes) {
            if (L.stream && L.stream != c->stream) cudaStreamDestroy(L.stream);
            if (L.done) cudaEventDestroy(L.done);
        }
        c->lanes.assign(nl, Lane());
        for (int i = 0; i < nl; ++i) {
            Lane& L = c->lanes[i];
            L.idx = i;
            L.f0 = (int)((long long)c->F * i / nl);
            L.F = (int)((long long)c->F * (i + 1) / nl) - L.f0;
            if (i == 0) L.stream = c->stream;
            else CK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
        }
        if (!c->evLaneFork) CK(cudaEventCreateWithFlags(&c->evLaneFork, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(c->evLaneFork, c->stream));
    for (Lane& L : c->lanes) {
        if (L.stream != c->stream) CK(cudaStreamWaitEvent(L.stream, c->evLaneFork, 0));
        L.v = lane_view(c, L.f0, L.F);
        L.hn.assign(c->hn.begin() + L.f0, c->hn.begin() + L.f0 + L.F);
        L.hN.assign(c->hN.begin() + L.f0, c->hN.begin() + L.f0 + L.F);
        L.hKp.assign(c->hKp.begin() + L.f0, c->hKp.begin() + L.f0 + L.F);
    }
    int rc;
    for (Lane& L : c->lanes) {
        LaneScope ls(c, &L);
        if ((rc = ekfb_predict(c)) != EKFB_OK) return rc;
        if ((rc = ekfb_measure(c)) != EKFB_OK) return rc;
        if ((rc = ekfb_match(c)) != EKFB_OK) return rc;
        L.chunk0 = 0;
        if ((rc = ransac_launch(c, 0, &L.seq)) != EKFB_OK) return rc;
        L.pending = true;
    }
    for (bool any = true; any;) {
        any = false;
        for (Lane& L : c->lanes) {
            if (!L.pending) continue;
            LaneScope ls(c, &L);
            bool done = false;
            if ((rc = ransac_collect(c, L.seq, L.chunk0, &done)) != EKFB_OK) return rc;
            if (!done) {
                L.chunk0 += ransac_chunk_len(c);
                if ((rc = ransac_launch(c, L.chunk0, &L.seq)) != EKFB_OK) return rc;
                any = true;
                continue;
            }
            L.pending = false;
            if ((rc = ekfb_update(c, 0)) != EKFB_OK) return rc;
            if ((rc = rescue_launch(c, &L.seq)) != EKFB_OK) return rc;
        }
    }
    for (Lane& L : c->lanes) {
        LaneScope ls(c, &L);
        if ((rc = wait_published_dims(c, L.seq)) != EKFB_OK) return rc;
        if ((rc = ekfb_update(c, 1)) != EKFB_OK) return rc;
        if (L.stream != ls.s0) CK(cudaEventRecord(L.done, L.stream));
    }
    for (Lane& L : c->lanes)
        if (L.stream != c->stream) CK(cudaStreamWaitEvent(c->stream, L.done, 0));
    return EKFB_OK;
}

extern "C" int ekfb_step(ekfb_handle c)
{
    int rc;
    REQUIRE(c, "null handle");
    struct InStep { ekfb_ctx* c; InStep(ekfb_ctx* c_) : c(c_) { c->in_step = true; } ~InStep() { c->in_step = false; } } inStep(c);
    const int nl = lanes_wanted(c);
    if (nl > 1) return step_lanes(c, nl);
    if ((rc = ekfb_predict(c)) != EKFB_OK) return rc;
    if ((rc = ekfb_measure(c)) != EKFB_OK) return rc;
    if ((rc = ekfb_match(c)) != EKFB_OK) return rc;
    if ((rc = ekfb_ransac(c)) != EKFB_OK) return rc;
    if ((rc = ekfb_update(c, 0)) != EKFB_OK) return rc;
    if ((rc = ekfb_rescue(c)) != EKFB_OK) return rc;
    if ((rc = ekfb_update(c, 1)) != EKFB_OK) return rc;
    return EKFB_OK;   // (the map-feature bookkeeping ran in the tail of the rescue pass)
}

// ---- results -----------------------------------------------------------------------------------
static void fill_info(const int* d, ekfb_frame_info* o)
{
    o->n = d[D_N_STATE]; o->n_features = d[D_N_FEAT]; o->n_keypoints = d[D_N_KP]; o->n_predicted = d[D_N_PRED];
    o->n_matches = d[D_N_MATCH]; o->n_hypotheses = d[D_N_HYP]; o->best_hypothesis = d[D_BEST_HYP];
    o->n_inliers = d[D_N_INL]; o->n_outliers = d[D_N_OUT]; o->n_rescued = d[D_N_RESC]; o->status = d[D_STATUS];
    o->reserved = d[D_STATUS_EVER] | d[D_STATUS];   /* any non-zero status since ekfb_set_state */
}

extern "C" int ekfb_get_frame_info(ekfb_handle c, int f, ekfb_frame_info* info)
{
    REQUIRE(c && info, "null argument");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    int rc = read_dims(c);
    if (rc != EKFB_OK) return rc;
    fill_info(c->h_dims + (size_t)f * D_STRIDE, info);
    return EKFB_OK;
}

static_assert(sizeof(RecordDev) == sizeof(ekfb_record), "record layouts must match");

extern "C" int ekfb_write_records_device(ekfb_handle c, void* device_out)
{
    REQUIRE(c && device_out, "null argument");
    CK(cudaSetDevice(c->device));
    k_write_records<<<c->F, 192, 0, c->stream>>>(c->v, reinterpret_cast<RecordDev*>(device_out));
    count_launch(c);
    CK(cudaGetLastError());
    return EKFB_OK;
}

extern "C" int ekfb_get_records(ekfb_handle c, ekfb_record* out)
{
    REQUIRE(c && out, "null argument");
    int rc = ekfb_write_records_device(c, c->d_rec);
    if (rc != EKFB_OK) return rc;
    CK(cudaMemcpyAsync(c->h_rec, c->d_rec, sizeof(RecordDev) * c->F, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(out, c->h_rec, sizeof(RecordDev) * c->F);
    return EKFB_OK;
}

#define D2H(dst, src, count, T)                                                                              \
    do {                                                                                                     \
        if (dst) CK(cudaMemcpyAsync(dst, (src), sizeof(T) * (size_t)(count), cudaMemcpyDeviceToHost, c->stream)); \
    } while (0)

extern "C" int ekfb_get_feature_results(ekfb_handle c, int f, uint8_t* predicted, double* hpred, double* S, double* Hx,
                                        double* Hf, uint8_t* matched, double* z, int32_t* kp_index, float* dist,
                                        uint8_t* inlier, uint8_t* outlier, uint8_t* rescued)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    const size_t fo = (size_t)f * c->Nmax;
    const int N = c->hN[f];
    DevView& v = c->v;
    D2H(predicted, v.vis + fo, N, uint8_t);
    D2H(hpred, v.h + fo * 2, N * 2, double);
    D2H(S, v.Si + fo * 4, N * 4, double);
    D2H(Hx, v.Hx + fo * 14, N * 14, double);
    D2H(Hf, v.Hf + fo * 12, N * 12, double);
    D2H(matched, v.mflag + fo, N, uint8_t);
    D2H(z, v.z + fo * 2, N * 2, double);
    D2H(kp_index, v.mkp + fo, N, int);
    D2H(dist, v.mdist + fo, N, float);
    D2H(inlier, v.inl + fo, N, uint8_t);
    D2H(outlier, v.outl + fo, N, uint8_t);
    D2H(rescued, v.resc + fo, N, uint8_t);
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_get_mask(ekfb_handle c, int f, uint8_t* mask, uint8_t* kp_ok)
{
    REQUIRE(c, "null handle");
    REQUIRE(f >= 0 && f < c->F, "bad filter index");
    CK(cudaSetDevice(c->device));
    DevView& v = c->v;
    D2H(mask, v.mask + (size_t)f * v.W * v.H, (size_t)v.W * v.H, uint8_t);
    D2H(kp_ok, v.kpok + (size_t)f * c->Kpmax, c->hKp[f], uint8_t);
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

// ---- isolated kernels -------------------------------------------------------------------------------
extern "C" int ekfb_test_downdate(ekfb_handle c, int n, int k, const double* P_in, const double* Wt, double* P_out)
{
    REQUIRE(c && P_in && Wt && P_out, "null argument");
    REQUIRE(n >= 13 && n <= c->nmax && k > 0 && k <= c->kmax && (k % 2) == 0, "n or k out of range (k must be even)");
    CK(cudaSetDevice(c->device));
    DevView& v = c->v;
    CK(cudaMemcpy2DAsync(v.P, sizeof(double) * c->ld, P_in, sizeof(double) * n, sizeof(double) * n, n,
                         cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(v.Bu, 0, sizeof(double) * (size_t)std::min(rup(k, 16), c->kmax) * c->ld, c->stream));
    CK(cudaMemcpy2DAsync(v.Bu, sizeof(double) * c->ld, Wt, sizeof(double) * n, sizeof(double) * n, k,
                         cudaMemcpyHostToDevice, c->stream));
    int* hd = c->h_dims;
    const int save_n = hd[D_N_STATE], save_u = hd[D_ULIST];
    hd[D_N_STATE] = n;
    hd[D_ULIST] = k / 2;
    CK(cudaMemcpyAsync(v.dims, hd, sizeof(int) * D_STRIDE, cudaMemcpyHostToDevice, c->stream));
    // once ekfb_flush_l2 has been used on this handle, the launch is measured with P and W evicted from L2
    if (c->flush_buf) CK(cudaMemsetAsync(c->flush_buf, 0, c->flush_bytes, c->stream));
    CK(cudaEventRecord(c->pe[0], c->stream));
    {
        const int saveN = c->hn[0];
        c->hn[0] = n;
        int rcD = launch_downdate(c, n);
        c->hn[0] = saveN;
        if (rcD != EKFB_OK) {   // leave the handle's counters as they were
            hd[D_N_STATE] = save_n;
            hd[D_ULIST] = save_u;
            cudaMemcpyAsync(v.dims, hd, sizeof(int) * D_STRIDE, cudaMemcpyHostToDevice, c->stream);
            cudaStreamSynchronize(c->stream);
            return rcD;
        }
    }
    CK(cudaEventRecord(c->pe[1], c->stream));
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(P_out, sizeof(double) * n, v.P, sizeof(double) * c->ld, sizeof(double) * n, n,
                         cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->last_downdate_ms, c->pe[0], c->pe[1]));
    hd[D_N_STATE] = save_n;
    hd[D_ULIST] = save_u;
    CK(cudaMemcpyAsync(v.dims, hd, sizeof(int) * D_STRIDE, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_test_factor(ekfb_handle c, int k, const double* S_in, double* U_out, double* Uinv_out)
{
    REQUIRE(c && S_in && U_out && Uinv_out, "null argument");
    REQUIRE(k > 0 && k <= c->kmax && (k % 2) == 0, "k out of range (k must be even)");
    CK(cudaSetDevice(c->device));
    DevView& v = c->v;
    const int nb = cdiv(k, kNB);
    CK(cudaMemcpy2DAsync(v.S, sizeof(double) * c->ldS, S_in, sizeof(double) * (k + 1), sizeof(double) * (k + 1), k,
                         cudaMemcpyHostToDevice, c->stream));
    int* hd = c->h_dims;
    const int save_u = hd[D_ULIST];
    hd[D_ULIST] = k / 2;
    hd[D_STATUS] = 0;
    CK(cudaMemcpyAsync(v.dims, hd, sizeof(int) * D_STRIDE, cudaMemcpyHostToDevice, c->stream));
    c->schain_eff = c->schain_variant >= 0 ? c->schain_variant : 0;
    int rc = launch_schain(c, k);
    if (rc != EKFB_OK) return rc;
    if (c->schain_eff >= 3) k_chain_finish<<<cdiv(c->F, 128), 128, 0, c->stream>>>(c->chainCtl, c->nbMax, c->F);
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(U_out, sizeof(double) * (k + 1), v.Sf, sizeof(double) * c->ldS, sizeof(double) * (k + 1), k,
                         cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(Uinv_out, v.Uinv, sizeof(double) * (size_t)nb * kNB * kNB, cudaMemcpyDeviceToHost, c->stream));
    int status = 0;
    CK(cudaMemcpyAsync(&status, v.dims + D_STATUS, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    hd[D_ULIST] = save_u;
    hd[D_STATUS] = 0;
    CK(cudaMemcpyAsync(v.dims, hd, sizeof(int) * D_STRIDE, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    REQUIRE(status == 0, "innovation covariance is not positive definite");
    return EKFB_OK;
}

extern "C" int ekfb_time_update(ekfb_handle c, int which, int reps, float* ms_total, float* ms_downdate)
{
    REQUIRE(c && reps > 0, "bad argument");
    CK(cudaSetDevice(c->device));
    // total: reps full updates on the current list; downdate: the P -= W W^T kernel alone, reps times
    CK(cudaEventRecord(c->timers[62], c->stream));
    for (int r = 0; r < reps; ++r) {
        int rc = run_update(c, which);
        if (rc != EKFB_OK) return rc;
    }
    CK(cudaEventRecord(c->timers[63], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->timers[62], c->timers[63]));
    if (ms_total) *ms_total = ms / reps;
    const int n = max_of(c->hn), nI = cdiv(n, kDTM);
    CK(cudaEventRecord(c->timers[62], c->stream));
    for (int r = 0; r < reps; ++r) k_downdate<<<dim3(nI * (nI + 1), c->F), 128, kDownSmemBytes, c->stream>>>(c->v);
    CK(cudaEventRecord(c->timers[63], c->stream));
    count_launch(c, reps);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&ms, c->timers[62], c->timers[63]));
    if (ms_downdate) *ms_downdate = ms / reps;
    return EKFB_OK;
}

// ---- timing ------------------------------------------------------------------------------------------
extern "C" int ekfb_timer_record(ekfb_handle c, int slot)
{
    REQUIRE(c && slot >= 0 && slot < 62, "bad timer slot");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->timers[slot], c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_timer_elapsed_ms(ekfb_handle c, int a, int b, float* ms)
{
    REQUIRE(c && ms && a >= 0 && a < 62 && b >= 0 && b < 62, "bad timer slot");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->timers[b]));
    CK(cudaEventElapsedTime(ms, c->timers[a], c->timers[b]));
    return EKFB_OK;
}

extern "C" int ekfb_profile_enable(ekfb_handle c, int on)
{
    REQUIRE(c, "null handle");
    c->prof = on != 0;
    std::memset(c->prof_ms, 0, sizeof(c->prof_ms));
    std::memset(c->prof_launch, 0, sizeof(c->prof_launch));
    return EKFB_OK;
}

extern "C" int ekfb_profile_read(ekfb_handle c, float* ms9, int32_t* launches9)
{
    REQUIRE(c, "null handle");
    for (int g = 0; g < G_COUNT; ++g) {
        if (ms9) ms9[g] = c->prof_ms[g];
        if (launches9) launches9[g] = c->prof_launch[g];
    }
    std::memset(c->prof_ms, 0, sizeof(c->prof_ms));
    std::memset(c->prof_launch, 0, sizeof(c->prof_launch));
    return EKFB_OK;
}

extern "C" int64_t ekfb_kernel_launches(ekfb_handle c) { return c ? c->launches : 0; }

extern "C" int ekfb_set_option(ekfb_handle c, int option, int value)
{
    REQUIRE(c, "null handle");
    REQUIRE(option >= EKFB_OPT_FORCE_GENERIC_FACTOR && option <= EKFB_OPT_SLAB_TRSM_MAX_K, "unknown option");
    if (option == EKFB_OPT_SMALL_UPDATE) { c->small_update = value; return EKFB_OK; }
    if (option == EKFB_OPT_LANES) { c->lanes_opt = value; return EKFB_OK; }
    if (option == 13) { c->dd_probe = value; return EKFB_OK; }
    if (option == EKFB_OPT_MATCHER) { c->matcher = value; return EKFB_OK; }
    if (option == EKFB_OPT_NCC_WARP) { c->ncc_warp = value; return EKFB_OK; }
    if (option == EKFB_OPT_NCC_TMA_WINDOW) { c->ncc_tma_window = value; return EKFB_OK; }
    if (option == EKFB_OPT_SLAB_TRSM_MAX_K) { c->slab_trsm_max_k = value; return EKFB_OK; }
    if (option == EKFB_OPT_DOWNDATE_CTAS) { c->dd_ctas_per_sm = value == 1 ? 1 : 2; return EKFB_OK; }
    if (option == EKFB_OPT_FAULT_INJECT) { c->v.faultInject = value; return EKFB_OK; }
    if (option == EKFB_OPT_DOWNDATE_SMALL_K) { c->downdate_small_k = value; return EKFB_OK; }
    if (option == EKFB_OPT_TRSM_STAGES) { c->trsm_stages = value; return EKFB_OK; }
    if (option == EKFB_OPT_TRSM_PAIR) { c->trsm_pair = value; return EKFB_OK; }
    if (option == EKFB_OPT_PDL) { c->use_pdl = value; return EKFB_OK; }
    if (option == EKFB_OPT_RANSAC_CHUNK) { c->ransac_chunk = value; return EKFB_OK; }
    if (option == EKFB_OPT_FORCE_GENERIC_FACTOR) c->force_generic = value;
    else if (option == EKFB_OPT_SCHAIN_VARIANT) c->schain_variant = value;
    else c->downdate_variant = value;
    return EKFB_OK;
}

extern "C" int ekfb_debug_read(ekfb_handle c, long long* out64)
{
    REQUIRE(c && out64, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out64, c->v.dbg, sizeof(long long) * 64, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_flush_l2(ekfb_handle c)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    if (!c->flush_buf) {
        c->flush_bytes = (size_t)256 << 20;  // > 126 MB L2
        CK(cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    CK(cudaMemsetAsync(c->flush_buf, 0, c->flush_bytes, c->stream));
    return EKFB_OK;
}

extern "C" int ekfb_downdate_timing(ekfb_handle c, int enable)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    if (enable && c->dd_ev.empty()) {
        c->dd_ev.resize(8192);
        for (cudaEvent_t& e : c->dd_ev) CK(cudaEventCreate(&e));
    }
    c->dd_timing = enable != 0;
    c->dd_used = 0;
    c->dd_flops = c->dd_bytes = 0.;
    c->dd_launch_flops.clear();
    return EKFB_OK;
}

// per-launch view of the same measurement (call before ekfb_downdate_stats, which resets it): duration [ms] and algorithmic
// flop n (n + 1) K summed over the filters of every timed downdate launch, oldest first
extern "C" int ekfb_downdate_launches(ekfb_handle c, int cap, float* ms, double* flops, int32_t* count)
{
    REQUIRE(c && count, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const int nl = (int)std::min<size_t>(c->dd_used / 2, c->dd_launch_flops.size());
    *count = nl;
    for (int i = 0; i < nl && i < cap; ++i) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, c->dd_ev[2 * i], c->dd_ev[2 * i + 1]));
        if (ms) ms[i] = t;
        if (flops) flops[i] = c->dd_launch_flops[i];
    }
    return EKFB_OK;
}

extern "C" int ekfb_downdate_stats(ekfb_handle c, double* ms_total, int64_t* launches, double* flops_total,
                                   double* bytes_min_total)
{
    REQUIRE(c, "null handle");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    double ms = 0.;
    for (size_t i = 0; i + 1 < c->dd_used; i += 2) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, c->dd_ev[i], c->dd_ev[i + 1]));
        ms += t;
    }
    if (ms_total) *ms_total = ms;
    if (launches) *launches = (int64_t)(c->dd_used / 2);
    if (flops_total) *flops_total = c->dd_flops;
    if (bytes_min_total) *bytes_min_total = c->dd_bytes;
    c->dd_used = 0;
    c->dd_flops = c->dd_bytes = 0.;
    c->dd_launch_flops.clear();
    return EKFB_OK;
}
