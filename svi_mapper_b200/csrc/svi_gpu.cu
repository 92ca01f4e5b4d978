// svi_gpu.cu -- libsvi_gpu.so: context, scratch management, kernel pipeline and the C-ABI of
// include/svi_gpu.h.  sm_100a only; there is no CPU fallback anywhere in this file.
//
// Execution model.  A batch of independent stereo pairs is cut into chunks of `chunk_frames`
// frames; chunk c runs on lane c % n_lanes.  A lane is one CUDA stream plus the scratch of one
// chunk (box-sum planes of both images, candidate lists, corner lists, bin lists); different lanes overlap
// each other's copies, wide kernels (Harris, match) and the one-CTA-per-frame selection kernel.
// Per-query entry points (triangulate, describe, track, landmark refinement) run on lane 0's stream.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/svi_gpu.h"
#include "brief_match.cuh"
#include "binned.cuh"
#include "harris.cuh"
#include "landmark_opt.cuh"
#include "select.cuh"
#include "track_plan.cuh"

using namespace svi;

namespace {

constexpr int kMaxLanes = 8;
constexpr int kSmallFrames = 2;          // calls of up to this many frames go through the pinned bounce buffer
constexpr size_t kOutBytesPerSlot = 8 + 8 + 24 + 32 + 32 + 4 + 4 + 1;   // uv_l uv_r xyz desc_l desc_r dist idx status
constexpr int kStages = 5;
const char* const kStageNames[kStages] = {"harris_box", "boxsum_right", "select_corners", "describe_left", "stereo_match"};
constexpr int kStageDescribe = 3;   // present only on the batch path (LEFT descriptors ahead of the matcher)
constexpr int kRawFactor = 4;   // the Harris kernel's list (local maxima above the TILE threshold) vs max_candidates

std::string g_create_error;

struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint16_t* box_l = nullptr;   // S(y,x) of LEFT
    uint16_t* box_r = nullptr;   // S(y,x) of RIGHT
    uint16_t* box_ls = nullptr;  // planes stored shifted by one element, S(y,x+1): the odd-aligned TMA copies
    uint16_t* box_rs = nullptr;
    uint32_t* frame_max = nullptr;
    int* cand_count = nullptr;
    unsigned long long* cand = nullptr;
    ushort2* det_xy = nullptr;
    ushort2* kp_xy = nullptr;
    int* n_det = nullptr;
    int* n_kp = nullptr;
    uint32_t* g_head = nullptr;
    uint32_t* g_next = nullptr;
    uint8_t* g_state = nullptr;
    // staging for the host-buffer entry points
    uint8_t* img_l = nullptr;
    uint8_t* img_r = nullptr;
    uint8_t* mask = nullptr;
    CUtensorMap map_l{}, map_r{}, map_ls{}, map_rs{};  // TMA views of the four planes: (W, chunk*H) u16, row pitch box_pitch
    CUtensorMap map_l_patch{};                          // LEFT plane again, box = the 56 x 49 patch of one descriptor
    // binned matcher (binned.cuh): key-points sorted by image bin, one tile of the planes per bin
    int* bin_start = nullptr;          // [chunk][n_bins + 1]
    uint32_t* bin_slot = nullptr;      // [chunk][max_corners]
    ushort2* bin_kp = nullptr;         // [chunk][max_corners]
    CUtensorMap map_l_bin{}, map_r_bin{}, map_rs_bin{};
    StereoOutDev out{};
    std::vector<cudaEvent_t> ev;  // stage boundary events (profiling)
    size_t ev_used = 0;
    std::vector<uint8_t> ev_pre;  // per profiled chunk: did describe_left_kernel run between its two events?
};

}  // namespace

struct svi_ctx {
    int device = 0;
    svi_camera cam_l{}, cam_r{};
    svi_params p{};
    int W = 0, H = 0;
    int dev_pitch = 0;   // bytes per row of the staged images
    int resp_pitch = 0, box_pitch = 0;
    int chunk = 0, n_lanes = 0;
    int cand_cap = 0, raw_cap = 0;
    float* resp_one = nullptr;   // svi_harris_response only: one W x H plane
    bool select_smem = true;
    int match_split = 0;     // warps per key-point in the scan-line matcher: 0 = chosen per launch; SVI_MATCH_SPLIT = 1 | 2 forces one
    bool trace = false;      // SVI_TRACE: host-side phase times of the tracking call on stderr (diagnostics)
    bool match_pre = true;   // LEFT descriptors by describe_left_kernel ahead of the matcher (SVI_MATCH_PRE = 0 | 1)
    bool match_binned = true;  // batch path: key-points binned by position, one TMA tile per bin (SVI_MATCH_BINNED = 0 | 1 | 2: off, by density, forced)
    int nbx = 0, n_bins = 0;
    SelectParams sel{};
    TriConst tc{};
    float f1 = 0, f0 = 0, kf = 0;
    Lane lanes[kMaxLanes];
    // candidate-list overflow flag: one int of mapped page-locked host memory.  Kernels write it in place (only when a
    // list overflows), the host reads it after the synchronisation it does anyway -- no copy, no extra round trip per call
    int* d_overflow = nullptr;            // device view
    volatile int* h_overflow = nullptr;   // host view
    cudaEvent_t fork = nullptr;
    // latency path (one pair per call): the RIGHT image travels and gets its box sums on a second stream while the
    // detector runs on LEFT; `side_done` joins it back before the first kernel that needs both
    cudaStream_t side = nullptr;
    cudaEvent_t side_done = nullptr;
    // tracking: both images of the current pair as two planes of one buffer, and the scratch of the
    // window-mode detector of stage 2 (grown on demand)
    uint8_t* trk_img = nullptr;
    struct RoiScratch {
        int items = 0;
        uint32_t* max = nullptr;
        int* cand_count = nullptr;
        unsigned long long* cand = nullptr;
        ushort2* det = nullptr;
        ushort2* kp = nullptr;
        int* n_det = nullptr;
        int* n_kp = nullptr;
        RoiItem* rois = nullptr;
        Stage2Item* s2 = nullptr;
        int* defer = nullptr;   // windows that do not fit the small selection configuration
        int* counters = nullptr;  // device-side item counts: stage 2 LEFT, stage 2 RIGHT, stage 3
    } roi;
    Stage3Item* s3_items = nullptr;
    // pinned bounce buffer of the small-call path (one or a few frames per call: the tracker's per-frame use)
    unsigned char* pin = nullptr;
    size_t pin_bytes = 0;
    // ... and its device-side result block: every output array of up to kSmallFrames frames in ONE allocation, so that the
    // results of a single-pair call come back with one copy
    unsigned char* small_block = nullptr;
    size_t small_bytes = 0;
    StereoOutDev small_out{};
    int* small_n_kp = nullptr;
    int* small_n_det = nullptr;
    // pinned mirror of the per-query arena: uploads and result downloads of the per-query / tracking entry points bounce
    // through it, so that pageable caller buffers never turn a transfer into a blocking staged copy
    unsigned char* pin_arena = nullptr;
    struct Pending { void* host; const unsigned char* pinned; size_t bytes; };
    std::vector<Pending> pending;
    int n_sm = 148;
    // svi_optimize_landmarks: one device buffer + its pinned mirror (inputs, then outputs), grown on demand
    unsigned char* opt_dev = nullptr;
    unsigned char* opt_pin = nullptr;
    size_t opt_bytes = 0;
    // per-query arena
    unsigned char* arena = nullptr;
    size_t arena_bytes = 0, arena_used = 0;
    bool profiling = false;
    bool batch_uploads = false;   // upload(): stage into the pinned mirror only (svi_track_landmarks sends one copy)
    bool serial = false;   // profiling mode 2: every chunk on lane 0, so that the stage events bracket ONE kernel each
    double stage_ms[kStages] = {};
    long stage_launches[kStages] = {};
    std::string err;
};

namespace {

int fail(svi_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(ctx, SVI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// grid of the Harris kernel: response tiles HT_SX x HT_SY apart, the tile interiors cover columns 1..W-2, rows 1..H-2
inline dim3 harris_grid(int W, int H, int layers) {
    return dim3((unsigned)std::max(1, (W - 2 + HT_SX - 1) / HT_SX), (unsigned)std::max(1, (H - 2 + HT_SY - 1) / HT_SY), (unsigned)layers);
}

FrameGeom make_geom(const svi_ctx* c, int img_pitch, size_t img_stride) {
    FrameGeom g;
    g.W = c->W; g.H = c->H;
    g.img_pitch = img_pitch;
    g.img_stride = img_stride;
    g.resp_pitch = c->resp_pitch;
    g.box_pitch = c->box_pitch;
    return g;
}

void* arena_alloc(svi_ctx* c, size_t bytes) {
    size_t off = (c->arena_used + 255) & ~size_t(255);
    if (off + bytes > c->arena_bytes) return nullptr;
    c->arena_used = off + bytes;
    return c->arena + off;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no libcuda link needed).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_box_map(CUtensorMap* m, uint16_t* base, int W, int rows, int box_pitch, std::string* err, int box_w = PATCH_W,
                  int box_h = PATCH_ROWS) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
            q != cudaDriverEntryPointSuccess) {
            *err = "cuTensorMapEncodeTiled is not available from this driver";
            return false;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)box_pitch * sizeof(uint16_t)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return false;
    }
    return true;
}

// Stage-boundary events come from a per-lane pool that svi_set_profiling fills ahead of time (kEventPool events per
// lane), so that no event is created inside a timed region; the pool still grows if a step needs more.
constexpr size_t kEventPool = 1024;
void reserve_events(Lane& l, size_t count) {
    while (l.ev.size() < count) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) break;
        l.ev.push_back(e);
    }
}
cudaEvent_t lane_event(Lane& l) {
    if (l.ev_used == l.ev.size()) reserve_events(l, l.ev.size() + 64);
    return l.ev[l.ev_used++];
}

void mark(svi_ctx* c, Lane& l) {
    if (c->profiling) cudaEventRecord(lane_event(l), l.stream);
}

// Fold the recorded stage events into per-stage totals (call after the lanes are idle).
void collect_timings(svi_ctx* c) {
    for (int li = 0; li < c->n_lanes; ++li) {
        Lane& l = c->lanes[li];
        for (size_t i = 0, rec = 0; i + kStages < l.ev_used; i += kStages + 1, ++rec) {
            for (int s = 0; s < kStages; ++s) {
                if (s == kStageDescribe && !(rec < l.ev_pre.size() && l.ev_pre[rec])) continue;   // no kernel in that bracket
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, l.ev[i + s], l.ev[i + s + 1]) == cudaSuccess) {
                    c->stage_ms[s] += ms;
                    c->stage_launches[s] += 1;
                }
            }
        }
        l.ev_used = 0;
        l.ev_pre.clear();
    }
}

// Descriptor + match stages of the batch path, binned form (binned.cuh): sort the key-points into bins, LEFT descriptors
// per bin, scan-line searches per bin.
template <class G>
int run_binned(svi_ctx* ctx, Lane& l, const FrameGeom& g, int nf, const StereoOutDev& out, int out_frame0, const int* n_kp) {
    cudaStream_t s = l.stream;
    const dim3 bgrid(ctx->n_bins, nf);
    bin_keypoints_kernel<<<nf, BINK_THREADS, sizeof(uint32_t) * ctx->n_bins, s>>>(l.kp_xy, n_kp, ctx->p.max_corners, G::BIN_W, G::BIN_H, ctx->nbx,
                                                                                 ctx->n_bins, l.bin_start, l.bin_slot, l.bin_kp);
    describe_left_binned_kernel<G><<<bgrid, G::D_WARPS * 32, G::D_SMEM, s>>>(l.map_l_bin, g, l.bin_start, l.bin_slot, l.bin_kp, ctx->nbx, ctx->n_bins,
                                                                            ctx->p.max_corners, out.desc_l, out.cap, out_frame0);
    if (ctx->profiling) l.ev_pre.push_back(1);
    mark(ctx, l);
    stereo_match_binned_kernel<G><<<bgrid, G::M_WARPS * 32, G::M_SMEM, s>>>(l.map_r_bin, l.map_rs_bin, g, ctx->tc, ctx->p.keypoint_size,
                                                                           ctx->p.search_range_px, l.bin_start, l.bin_slot, l.bin_kp, ctx->nbx,
                                                                           ctx->n_bins, ctx->p.max_corners, out, out_frame0, ctx->d_overflow);
    mark(ctx, l);
    return SVI_SUCCESS;
}

template <class G>
int setup_binned(svi_ctx* ctx, Lane& l, size_t rows, std::string* err) {
    if (!make_box_map(&l.map_l_bin, l.box_l, ctx->W, (int)rows, ctx->box_pitch, err, G::D_COLS, G::ROWS) ||
        !make_box_map(&l.map_r_bin, l.box_r, ctx->W, (int)rows, ctx->box_pitch, err, G::COLS, G::ROWS) ||
        !make_box_map(&l.map_rs_bin, l.box_rs, ctx->W, (int)rows, ctx->box_pitch, err, G::COLS, G::ROWS))
        return SVI_ERR_CUDA;
    return SVI_SUCCESS;
}

// The kernels of the new-landmark path for `nf` frames on one lane: detector, RIGHT box sums, corner selection, then the
// descriptor / match stages in the form that fits the call (svi_kernels_per_chunk: 4 to 6 launches).
int run_pipeline(svi_ctx* ctx, Lane& l, const uint8_t* d_left, const uint8_t* d_right, const uint8_t* d_mask,
                 const FrameGeom& g, int nf, const StereoOutDev& out, int out_frame0, int* n_kp, int* n_det,
                 cudaStream_t side = nullptr, const std::function<int(cudaStream_t)>* upload_right = nullptr) {
    cudaStream_t s = l.stream;
    CK(cudaMemsetAsync(l.frame_max, 0, sizeof(uint32_t) * nf, s));
    CK(cudaMemsetAsync(l.cand_count, 0, sizeof(int) * nf, s));
    const dim3 tiles((g.W + HT_W - 1) / HT_W, (g.H + HT_H - 1) / HT_H, nf);
    const bool fast = ctx->p.detector == SVI_DETECTOR_FAST_9_16;
    mark(ctx, l);
    if (fast) {
        fast_candidates_kernel<<<tiles, HT_THREADS, 0, s>>>(d_left, d_mask, g, ctx->p.fast_threshold, ctx->p.fast_nonmax, l.cand,
                                                            l.cand_count, ctx->raw_cap);
        boxsum9_kernel<<<tiles, HT_THREADS, 0, s>>>(d_left, g, l.box_l, nullptr);
    } else {
        harris_box_kernel<<<harris_grid(g.W, g.H, nf), HT_THREADS, sizeof(HarrisSmem), s>>>(
            d_left, d_mask, g, ctx->f1, ctx->f0, ctx->kf, ctx->p.quality_level, nullptr, l.box_l, nullptr, l.frame_max, l.cand,
            l.cand_count, ctx->raw_cap, nullptr, nullptr, g.H);
    }
    mark(ctx, l);
    // `side` (latency path): the RIGHT plane goes up on that stream -- staged by the caller's hook only now, so that the
    // host-side copy into the pinned mirror overlaps the detector -- and its box sums run there too
    if (upload_right) {
        const int rc = (*upload_right)(side ? side : s);
        if (rc != SVI_SUCCESS) return rc;
    }
    boxsum9_kernel<<<tiles, HT_THREADS, 0, side ? side : s>>>(d_right, g, l.box_r, l.box_rs);
    if (side) {
        CK(cudaEventRecord(ctx->side_done, side));
        CK(cudaStreamWaitEvent(s, ctx->side_done, 0));   // the selection kernel below does not need it, the matcher does;
                                                          // one wait here keeps the stage events of the profiler in order
    }
    mark(ctx, l);
    const uint32_t* fmax = fast ? nullptr : l.frame_max;
    if (ctx->select_smem) {
        select_corners_kernel<true><<<nf, SEL_THREADS, 13 * SEL_SMEM_KEYS, s>>>(
            l.cand, l.cand_count, fmax, ctx->sel, nullptr, nullptr, nullptr, l.det_xy, n_det, l.kp_xy, n_kp, ctx->d_overflow, nullptr);
    } else {
        select_corners_kernel<false><<<nf, SEL_THREADS, 4 * SEL_SMEM_CELLS, s>>>(
            l.cand, l.cand_count, fmax, ctx->sel, l.g_head, l.g_next, l.g_state, l.det_xy, n_det, l.kp_xy, n_kp, ctx->d_overflow, nullptr);
    }
    mark(ctx, l);
    // full batches amortise a warp's set-up over MATCH_KP_PER_WARP key-points; a small call (one pair per frame in a
    // tracker) would leave most SMs idle that way, so it spreads the key-points until two waves of warps exist
    const long long slots = (long long)nf * ctx->p.max_corners;
    const int kp_per_warp = (int)std::max<long long>(1, std::min<long long>(MATCH_KP_PER_WARP, slots / (2LL * ctx->n_sm * 9)));
    const int kp_per_cta = MATCH_WARPS * kp_per_warp;
    const dim3 mgrid((ctx->p.max_corners + kp_per_cta - 1) / kp_per_cta, nf);
    // full batches: the LEFT descriptors first (one TMA patch per key-point), then the matcher with match_split warps per
    // key-point; small calls keep the single fused kernel (one launch less on the latency path)
    const bool pre = ctx->match_pre && kp_per_warp == MATCH_KP_PER_WARP;
    // warps per key-point: measured -- two warps win whenever the LEFT gathers stay in the matcher
    // (small calls) and on multi-pass searches (scan lines longer than one window), one warp wins on the batch path
    const int split = ctx->match_split ? ctx->match_split : ((pre && ctx->p.search_range_px <= (float)PATCH_CHUNK) ? 1 : 2);
    // batch path with the reference's geometry (scan line of one pass, whole-pixel ROI border): key-points binned by
    // position, one shared tile per bin (binned.cuh)
    if (pre && ctx->match_binned && l.bin_start) {
        const int rc = run_binned<BinWide>(ctx, l, g, nf, out, out_frame0, n_kp);
        CK(cudaGetLastError());
        return rc;
    }
    if (pre) {
        const dim3 dgrid((ctx->p.max_corners + DL_WARPS * DL_KP_PER_WARP - 1) / (DL_WARPS * DL_KP_PER_WARP), nf);
        describe_left_kernel<<<dgrid, DL_WARPS * 32, DL_SMEM, s>>>(l.map_l_patch, g, l.kp_xy, n_kp, ctx->p.max_corners, out.desc_l, out.cap,
                                                                   out_frame0);
    }
    if (ctx->profiling) l.ev_pre.push_back(pre ? 1 : 0);
    mark(ctx, l);
#define SVI_MATCH_ARGS l.box_l, l.map_r, l.map_rs, g, ctx->tc, ctx->p.keypoint_size, ctx->p.search_range_px, l.kp_xy, n_kp, \
                       ctx->p.max_corners, out, out_frame0, kp_per_warp
    if (pre && split == 1)
        stereo_match_split_kernel<1, true><<<mgrid, MatchSplit<1>::THREADS, MatchSplit<1>::SMEM, s>>>(SVI_MATCH_ARGS);
    else if (pre && split == 2)
        stereo_match_split_kernel<2, true><<<mgrid, MatchSplit<2>::THREADS, MatchSplit<2>::SMEM, s>>>(SVI_MATCH_ARGS);
    else if (split == 2)
        stereo_match_split_kernel<2, false><<<mgrid, MatchSplit<2>::THREADS, MatchSplit<2>::SMEM, s>>>(SVI_MATCH_ARGS);
    else
        stereo_match_kernel<<<mgrid, MATCH_WARPS * 32, MATCH_SMEM, s>>>(SVI_MATCH_ARGS);
#undef SVI_MATCH_ARGS
    mark(ctx, l);
    CK(cudaGetLastError());
    return SVI_SUCCESS;
}

int check_overflow(svi_ctx* ctx) {
#ifdef SVI_BOUNDS_CHECK
    {
        int line = 0;
        CK(cudaMemcpyFromSymbol(&line, g_svi_check, sizeof(int)));
        if (line) return fail(ctx, SVI_ERR_CUDA, "bounds check failed in a kernel: tag/line " + std::to_string(line));
    }
#endif
    const int h = *ctx->h_overflow;   // every caller has synchronised the streams that could have written it
    if (h) *ctx->h_overflow = 0;
    if (h == 4) return fail(ctx, SVI_ERR_CUDA, "internal error: a search window left its tile in the binned matcher (SVI_MATCH_BINNED=0 selects the per-key-point kernels)");
    if (h == 3) {
        return fail(ctx, SVI_ERR_CAPACITY, "FAST found more corners than svi_params.max_corners in a frame (cv::FAST returns all of them): raise max_corners");
    }
    if (h) {
        return fail(ctx, SVI_ERR_CAPACITY,
                    "NMS candidate list overflow: raise svi_params.max_candidates (a frame produced more than " +
                        std::to_string(ctx->cand_cap) + " candidates)");
    }
    return SVI_SUCCESS;
}

int sync_lanes(svi_ctx* ctx) {
    for (int i = 0; i < ctx->n_lanes; ++i) CK(cudaStreamSynchronize(ctx->lanes[i].stream));
    if (ctx->profiling) collect_timings(ctx);
    return SVI_SUCCESS;
}

template <typename T>
cudaError_t dmalloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)); }

}  // namespace

// ---- per-query entry points: one image (or pair) staged in lane 0, queries in the arena ----
namespace {
// Page-locked caller memory (cudaMallocHost, cudaHostRegister, a pinned torch tensor: a camera driver's DMA buffers) is
// read by the copy engine directly; anything else goes through the library's pinned mirror, because a cudaMemcpyAsync
// from pageable memory is a blocking staged copy.  (One driver query per image, ~1 us; the mirror memcpy of a
// 1241 x 376 pair costs ~75 us of a 0.2 ms single-pair call.)
bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int stage_box(svi_ctx* ctx, const uint8_t* img, size_t pitch, uint8_t* d_img, uint16_t* d_box, uint16_t* d_box_shift,
              cudaStream_t s, int pin_plane = 0) {
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, (size_t)ctx->H * ctx->dev_pitch);
    const size_t plane = (size_t)ctx->H * ctx->dev_pitch;
    if (host_is_pinned(img)) {
        if ((int)pitch == ctx->dev_pitch) CK(cudaMemcpyAsync(d_img, img, plane, cudaMemcpyHostToDevice, s));
        else CK(cudaMemcpy2DAsync(d_img, ctx->dev_pitch, img, pitch, ctx->W, ctx->H, cudaMemcpyHostToDevice, s));
    } else if (ctx->pin && pin_plane >= 0 && (size_t)(pin_plane + 1) * plane <= ctx->pin_bytes) {
        unsigned char* stage = ctx->pin + (size_t)pin_plane * plane;
        if ((int)pitch == ctx->dev_pitch) std::memcpy(stage, img, plane);
        else for (int y = 0; y < ctx->H; ++y) std::memcpy(stage + (size_t)y * ctx->dev_pitch, img + (size_t)y * pitch, ctx->W);
        CK(cudaMemcpyAsync(d_img, stage, plane, cudaMemcpyHostToDevice, s));
    } else {
        CK(cudaMemcpy2DAsync(d_img, ctx->dev_pitch, img, pitch, ctx->W, ctx->H, cudaMemcpyHostToDevice, s));
    }
    const dim3 tiles((g.W + HT_W - 1) / HT_W, (g.H + HT_H - 1) / HT_H, 1);
    boxsum9_kernel<<<tiles, HT_THREADS, 0, s>>>(d_img, g, d_box, d_box_shift);
    CK(cudaGetLastError());
    return SVI_SUCCESS;
}
template <typename T>
int upload(svi_ctx* ctx, T** d, const T* h, size_t count, cudaStream_t s) {
    *d = static_cast<T*>(arena_alloc(ctx, count * sizeof(T)));
    if (!*d) return fail(ctx, SVI_ERR_CAPACITY, "query arena exhausted: raise svi_params.max_queries");
    if (h && ctx->pin_arena) {
        unsigned char* mirror = ctx->pin_arena + (reinterpret_cast<unsigned char*>(*d) - ctx->arena);
        std::memcpy(mirror, h, count * sizeof(T));
        // batch_uploads: the caller sends the whole staged range with ONE copy once every input is in the mirror
        if (!ctx->batch_uploads) CK(cudaMemcpyAsync(*d, mirror, count * sizeof(T), cudaMemcpyHostToDevice, s));
    } else if (h) {
        CK(cudaMemcpyAsync(*d, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    return SVI_SUCCESS;
}
// Result array that lives in the arena -> caller's buffer: async copy into the pinned mirror now, memcpy after the sync.
int download(svi_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t s) {
    const unsigned char* d = static_cast<const unsigned char*>(dev);
    if (ctx->pin_arena && d >= ctx->arena && d + bytes <= ctx->arena + ctx->arena_bytes) {
        unsigned char* mirror = ctx->pin_arena + (d - ctx->arena);
        CK(cudaMemcpyAsync(mirror, dev, bytes, cudaMemcpyDeviceToHost, s));
        ctx->pending.push_back({host, mirror, bytes});
    } else {
        CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s));
    }
    return SVI_SUCCESS;
}
int flush_downloads(svi_ctx* ctx, cudaStream_t s) {
    CK(cudaStreamSynchronize(s));
    for (const svi_ctx::Pending& q : ctx->pending) std::memcpy(q.host, q.pinned, q.bytes);
    ctx->pending.clear();
    return SVI_SUCCESS;
}
#define UP(d, h, count)                                        \
    do {                                                       \
        int rc_ = upload(ctx, &(d), (h), (count), s);          \
        if (rc_ != SVI_SUCCESS) return rc_;                    \
    } while (0)
}  // namespace


namespace {
// getMaskActiveLandmarks on the device: lane.mask (plane 0) = 255 with a radius-7 disc of zeros per centre.
int build_mask(svi_ctx* ctx, Lane& l, const float* centres, int n_centres) {
    cudaStream_t s = l.stream;
    const size_t bytes = (size_t)ctx->H * ctx->dev_pitch;
    mask_fill_kernel<<<(unsigned)((bytes / 16 + 256) / 256), 256, 0, s>>>(l.mask, bytes);
    if (n_centres > 0) {
        ctx->arena_used = 0; ctx->pending.clear();
        float* d_c;
        UP(d_c, centres, (size_t)n_centres * 2);
        mask_discs_kernel<<<(n_centres * 15 + 255) / 256, 256, 0, s>>>(l.mask, ctx->W, ctx->H, ctx->dev_pitch, d_c, n_centres);
    }
    CK(cudaGetLastError());
    return SVI_SUCCESS;
}

// ---- small call (the tracker's one pair per frame).  Pageable host buffers make every cudaMemcpyAsync a blocking
// staged copy; here the images are packed into the pinned buffer, every transfer is a real async DMA on the lane's
// stream, and the outputs come back as one batch that is scattered with memcpy.  The detection mask is either the
// caller's plane(s) or -- with `centres` -- built on the device from the landmark centres
// (getMaskActiveLandmarks, CFundamentalMatcher.cpp:2043-2073): 8 bytes per landmark go up instead of a W x H plane.
int small_call(svi_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t pitch, size_t frame_stride, int n_frames,
               const uint8_t* masks, const float* centres, int n_centres, svi_stereo_result* out) {
    const int W = ctx->W, H = ctx->H, MC = ctx->p.max_corners, cap = out->capacity_per_frame;
    const size_t dstride = (size_t)H * ctx->dev_pitch;
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, dstride);
    {
        Lane& l = ctx->lanes[0];
        cudaStream_t s = l.stream;
        const int nf = n_frames;
        const size_t plane = dstride * nf;
        unsigned char* pin_in = ctx->pin;
        unsigned char* pin_out = ctx->pin + ((3 * (size_t)kSmallFrames * dstride + 15) & ~size_t(15));
        // upload order LEFT, mask, RIGHT: the detector starts as soon as LEFT (and the mask) is there; RIGHT goes up on the
        // side stream and its box sums run beside the detector (not while profiling: stage events live on one stream)
        cudaStream_t side = ctx->profiling ? nullptr : ctx->side;
        if (side) {   // whatever is still queued on the lane (an un-synchronised device-resident batch) uses the same scratch
            CK(cudaEventRecord(ctx->fork, s));
            CK(cudaStreamWaitEvent(side, ctx->fork, 0));
        }
        const uint8_t* srcs[3] = {left, masks, right};
        uint8_t* dsts[3] = {l.img_l, l.mask, l.img_r};
        auto upload_plane = [&](int k, cudaStream_t sk) -> int {
            if (!srcs[k]) return SVI_SUCCESS;
            if (host_is_pinned(srcs[k])) {   // the caller's buffer is page-locked: DMA straight from it
                if ((int)pitch == ctx->dev_pitch && (nf == 1 || frame_stride == dstride)) {
                    CK(cudaMemcpyAsync(dsts[k], srcs[k], plane, cudaMemcpyHostToDevice, sk));
                } else {
                    for (int f = 0; f < nf; ++f)
                        CK(cudaMemcpy2DAsync(dsts[k] + f * dstride, ctx->dev_pitch, srcs[k] + (size_t)f * frame_stride, pitch, W, H,
                                             cudaMemcpyHostToDevice, sk));
                }
                return SVI_SUCCESS;
            }
            unsigned char* stage = pin_in + k * (size_t)kSmallFrames * dstride;
            for (int f = 0; f < nf; ++f) {
                const uint8_t* src = srcs[k] + (size_t)f * frame_stride;
                if ((int)pitch == ctx->dev_pitch) std::memcpy(stage + f * dstride, src, dstride);
                else for (int y = 0; y < H; ++y) std::memcpy(stage + f * dstride + (size_t)y * ctx->dev_pitch, src + (size_t)y * pitch, W);
            }
            CK(cudaMemcpyAsync(dsts[k], stage, plane, cudaMemcpyHostToDevice, sk));
            return SVI_SUCCESS;
        };
        for (int k = 0; k < 2; ++k) {
            const int rc = upload_plane(k, s);
            if (rc != SVI_SUCCESS) return rc;
        }
        const std::function<int(cudaStream_t)> upload_right = [&](cudaStream_t sk) { return upload_plane(2, sk); };
        bool have_mask = masks != nullptr;
        if (centres) {
            int rc = build_mask(ctx, l, centres, n_centres);
            if (rc != SVI_SUCCESS) return rc;
            have_mask = true;
        }
        int rc = run_pipeline(ctx, l, l.img_l, l.img_r, have_mask ? l.mask : nullptr, g, nf, ctx->small_out, 0, ctx->small_n_kp, ctx->small_n_det, side,
                              &upload_right);
        if (rc != SVI_SUCCESS) { cudaStreamSynchronize(s); if (side) cudaStreamSynchronize(side); return rc; }
        // one copy brings the whole result block back; only the live slots are scattered to the caller's arrays
        CK(cudaMemcpyAsync(pin_out, ctx->small_block, ctx->small_bytes, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const StereoOutDev& so = ctx->small_out;
        auto host = [&](const void* d) { return pin_out + (static_cast<const unsigned char*>(d) - ctx->small_block); };
        const int* pin_kp = reinterpret_cast<const int*>(host(ctx->small_n_kp));
        const int* pin_det = reinterpret_cast<const int*>(host(ctx->small_n_det));
        struct Part { const void* dev; void* host; size_t elem; };   // elem = bytes per key-point slot
        const Part parts[8] = {{so.uv_l, out->uv_left, 8}, {so.uv_r, out->uv_right, 8}, {so.xyz, out->xyz_left, 24},
                               {so.desc_l, out->desc_left, 32}, {so.desc_r, out->desc_right, 32}, {so.dist, out->distance, 4},
                               {so.idx, out->match_index, 4}, {so.status, out->status, 1}};
        for (const Part& q : parts)
            for (int f = 0; f < nf; ++f)
                std::memcpy(static_cast<unsigned char*>(q.host) + (size_t)f * cap * q.elem, host(q.dev) + (size_t)f * MC * q.elem,
                            (size_t)std::min(std::max(pin_kp[f], 0), MC) * q.elem);
        std::memcpy(out->n_keypoints, pin_kp, sizeof(int) * nf);
        if (out->n_detected) std::memcpy(out->n_detected, pin_det, sizeof(int) * nf);
        return check_overflow(ctx);
    }
}

int triangulate_common(svi_ctx* ctx, bool left_search, const uint8_t* img, size_t pitch, int n, const float* search_range,
                       const float* top_left, const float* uv_ref, const uint8_t* desc_ref, float size, svi_tri_result* out) {
    if (!img || !top_left || !uv_ref || !desc_ref || !out || n < 0 || (int)pitch < ctx->W || (left_search && !search_range))
        return fail(ctx, SVI_ERR_INVALID, "svi_triangulate: bad argument");
    if (!out->uv || !out->xyz_left || !out->desc || !out->distance || !out->match_index || !out->status)
        return fail(ctx, SVI_ERR_INVALID, "svi_triangulate: null output array");
    if (n > ctx->p.max_queries) return fail(ctx, SVI_ERR_CAPACITY, "svi_triangulate: n > max_queries");
    if (n == 0) return SVI_SUCCESS;
    CK(cudaSetDevice(ctx->device));
    Lane& l = ctx->lanes[0];
    cudaStream_t s = l.stream;
    ctx->arena_used = 0; ctx->pending.clear();
    int rc = stage_box(ctx, img, pitch, l.img_l, l.box_l, l.box_ls, s);
    if (rc != SVI_SUCCESS) return rc;
    float* d_range = nullptr; float* d_tl; float* d_uv; uint8_t* d_desc;
    TriOutDev o;
    if (left_search) UP(d_range, search_range, (size_t)n);
    UP(d_tl, top_left, (size_t)n * 2);
    UP(d_uv, uv_ref, (size_t)n * 2);
    UP(d_desc, desc_ref, (size_t)n * 32);
    UP(o.uv, (const float*)nullptr, (size_t)n * 2);
    UP(o.xyz, (const double*)nullptr, (size_t)n * 3);
    UP(o.desc, (const uint8_t*)nullptr, (size_t)n * 32);
    UP(o.dist, (const int*)nullptr, (size_t)n);
    UP(o.idx, (const int*)nullptr, (size_t)n);
    UP(o.status, (const uint8_t*)nullptr, (size_t)n);
    CK(cudaMemsetAsync(o.uv, 0, sizeof(float) * 2 * n, s));
    CK(cudaMemsetAsync(o.xyz, 0, sizeof(double) * 3 * n, s));
    CK(cudaMemsetAsync(o.desc, 0, (size_t)32 * n, s));
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, (size_t)ctx->H * ctx->dev_pitch);
    const int blocks = (n + MATCH_WARPS - 1) / MATCH_WARPS;
    if (left_search)
        triangulate_kernel<true><<<blocks, MATCH_WARPS * 32, MATCH_SMEM, s>>>(l.map_l, l.map_ls, g, ctx->tc, n, d_range, d_tl, d_uv, d_desc, size, o);
    else
        triangulate_kernel<false><<<blocks, MATCH_WARPS * 32, MATCH_SMEM, s>>>(l.map_l, l.map_ls, g, ctx->tc, n, nullptr, d_tl, d_uv, d_desc, size, o);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out->uv, o.uv, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->xyz_left, o.xyz, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->desc, o.desc, (size_t)32 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->distance, o.dist, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->match_index, o.idx, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out->status, o.status, (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SVI_SUCCESS;
}
}  // namespace



namespace {

constexpr int kRoiRawCap = 8192;   // local maxima per stage-2 window (windows are at most ~600 x 600 px at the largest scaling)
constexpr int kRoiBatch = 4096;    // landmarks per wave of the window-mode detector

// Scratch of the tracking cascade: allocated once, at the first tracking call, sized from max_queries; never
// reallocated or freed inside a call.
int ensure_track_scratch(svi_ctx* ctx) {
    svi_ctx::RoiScratch& r = ctx->roi;
    if (r.items) return SVI_SUCCESS;
    const int items = std::min(std::max(ctx->p.max_queries, 64), kRoiBatch);
    const size_t MC = (size_t)ctx->p.max_corners;
    // [max | cand_count | defer | counters]: one block, so that a stage-2 side zeroes its bookkeeping with one memset
    CK(dmalloc(&r.max, (size_t)items * 3 + 4));
    r.cand_count = reinterpret_cast<int*>(r.max) + items;
    r.defer = r.cand_count + items;
    r.counters = r.defer + items;
    CK(dmalloc(&r.cand, (size_t)items * kRoiRawCap));
    CK(dmalloc(&r.det, (size_t)items * MC));
    CK(dmalloc(&r.kp, (size_t)items * MC));
    CK(dmalloc(&r.n_det, (size_t)items));
    CK(dmalloc(&r.n_kp, (size_t)items));
    CK(dmalloc(&r.rois, (size_t)items));
    CK(dmalloc(&r.s2, (size_t)items));
    CK(dmalloc(&ctx->s3_items, (size_t)std::max(ctx->p.max_queries, 64)));
    r.items = items;
    return SVI_SUCCESS;
}

TrackPlanConst make_plan(const svi_ctx* ctx, const double* T, const svi_camera& cam, double motion_scaling) {
    TrackPlanConst k;
    for (int i = 0; i < 12; ++i) { k.T[i] = T[i]; k.P[i] = cam.P[i]; }
    k.motion_scaling = motion_scaling;
    k.W = ctx->W; k.H = ctx->H;
    k.block = 15;        // m_uSearchBlockSizePoseOptimization (CFundamentalMatcher.h:95)
    k.epi_base = 15.0;   // m_dEpipolarLineBaseLength (CFundamentalMatcher.h:92)
    return k;
}

// One side of trackManual stage 2 for every landmark that is still untracked, entirely enqueued on the lane's stream:
// window arithmetic (:1548-1575) in stage2_plan_kernel, GFTT inside the windows (window-mode Harris + selection),
// description, matching and triangulation in track_stage2_kernel.  Landmarks go through in waves of kRoiBatch.
int track_stage2_side(svi_ctx* ctx, Lane& l, const FrameGeom& g, int n, const double* T, double motion_scaling, bool left,
                      const LandmarksDev& ld, const TrackOutDev& o) {
    cudaStream_t s = l.stream;
    const svi_camera& cam = left ? ctx->cam_l : ctx->cam_r;
    svi_ctx::RoiScratch& r = ctx->roi;
    if (!ctx->select_smem) return fail(ctx, SVI_ERR_UNSUPPORTED, "svi_track_landmarks: stage 2 needs max_candidates <= 16384");
    const TrackPlanConst k = make_plan(ctx, T, cam, motion_scaling);
    // upper bound of the window size (the grid of the window-mode detector): the principal weight is largest at the
    // image edge that is farthest from the principal point
    const double cx = cam.P[2], cy = cam.P[6];
    const double su = std::round(std::sqrt(std::max(std::fabs(cx), std::fabs((double)ctx->W - cx))) / 10.0 + motion_scaling);
    const double sv = std::round(std::sqrt(std::max(std::fabs(cy), std::fabs((double)ctx->H - cy))) / 10.0 + motion_scaling);
    const int max_w = std::min(ctx->W, 2 * (int)std::round(su * 15.0) + 2), max_h = std::min(ctx->H, 2 * (int)std::round(sv * 15.0) + 2);
    if (max_w <= 0 || max_h <= 0) return SVI_SUCCESS;   // no window can exist: the plan kernel reports "no features" per landmark
    SelectParams sp = ctx->sel;
    sp.raw_cap = kRoiRawCap;
    sp.cand_cap = std::min(ctx->cand_cap, kRoiRawCap);
    sp.cell = std::max(1, (int)std::ceil(ctx->p.min_distance));
    sp.cell_magic = sp.cell < 256 ? (uint32_t)(((1u << 24) + sp.cell - 1) / sp.cell) : 0u;
    int* n_items = r.counters + (left ? 0 : 1);
    for (int q0 = 0; q0 < n; q0 += r.items) {
        const int nb = std::min(r.items, n - q0);
        CK(cudaMemsetAsync(r.max, 0, sizeof(int) * ((size_t)r.items * 3 + 4), s));   // max, cand_count, defer, counters
        stage2_plan_kernel<<<(nb + 127) / 128, 128, 0, s>>>(k, left ? 0 : 1, (float)(1.0 + motion_scaling), ld, q0, q0 + nb, o, r.rois, r.s2, n_items);
        harris_box_kernel<<<harris_grid(max_w, max_h, nb), HT_THREADS, sizeof(HarrisSmem), s>>>(
            ctx->trk_img, nullptr, g, ctx->f1, ctx->f0, ctx->kf, ctx->p.quality_level, nullptr, nullptr, nullptr, r.max, r.cand,
            r.cand_count, kRoiRawCap, r.rois, n_items, 0);
        // thousands of small windows: seven small-configuration CTAs per SM; the rare window that does not fit is
        // deferred to the frame-size configuration (its CTAs return at once for every other window)
        select_corners_kernel<true, SEL_SMALL_THREADS, SEL_SMALL_KEYS, SEL_SMALL_CELLS>
            <<<nb, SEL_SMALL_THREADS, select_smem_bytes(SEL_SMALL_KEYS, SEL_SMALL_CELLS), s>>>(
                r.cand, r.cand_count, r.max, sp, nullptr, nullptr, nullptr, r.det, r.n_det, r.kp, r.n_kp, ctx->d_overflow, r.rois, r.defer, nullptr, n_items);
        select_corners_kernel<true><<<nb, SEL_THREADS, 13 * SEL_SMEM_KEYS, s>>>(r.cand, r.cand_count, r.max, sp, nullptr, nullptr, nullptr, r.det,
                                                                              r.n_det, r.kp, r.n_kp, ctx->d_overflow, r.rois, nullptr, r.defer, n_items);
        const int blocks = (nb + MATCH_WARPS - 1) / MATCH_WARPS;
        if (left)
            track_stage2_kernel<true><<<blocks, MATCH_WARPS * 32, MATCH_SMEM, s>>>(l.box_l, l.map_r, l.map_rs, g, ctx->tc, ctx->p.cutoff_stage2,
                                                                                  r.s2, n_items, r.det, r.n_det, ctx->p.max_corners, ld, o);
        else
            track_stage2_kernel<false><<<blocks, MATCH_WARPS * 32, MATCH_SMEM, s>>>(l.box_r, l.map_l, l.map_ls, g, ctx->tc, ctx->p.cutoff_stage2,
                                                                                   r.s2, n_items, r.det, r.n_det, ctx->p.max_corners, ld, o);
        CK(cudaGetLastError());
    }
    return SVI_SUCCESS;
}

// Stage 3 (epipolar line in LEFT): line geometry per landmark in stage3_plan_kernel (:1795-1947), sampling / matching /
// triangulation in track_stage3_kernel, both enqueued on the lane's stream.
int track_stage3_all(svi_ctx* ctx, Lane& l, const FrameGeom& g, int n, const double* T, double motion_scaling, const LandmarksDev& ld,
                     const Stage3Extra& ex, const uint8_t* d_orig, const TrackOutDev& o) {
    cudaStream_t s = l.stream;
    int* n_items = ctx->roi.counters + 2;
    const TrackPlanConst k = make_plan(ctx, T, ctx->cam_l, motion_scaling);
    CK(cudaMemsetAsync(n_items, 0, sizeof(int), s));
    stage3_plan_kernel<<<(n + 127) / 128, 128, 0, s>>>(k, ld, ex, n, o, ctx->s3_items, n_items);
    track_stage3_kernel<<<(n + MATCH_WARPS - 1) / MATCH_WARPS, MATCH_WARPS * 32, MATCH_SMEM, s>>>(
        l.box_l, l.map_r, l.map_rs, g, ctx->tc, ctx->p.cutoff_stage3, ctx->p.cutoff_original, ctx->s3_items, n_items, d_orig, ld, o);
    CK(cudaGetLastError());
    return SVI_SUCCESS;
}

}  // namespace

extern "C" {

int svi_params_default(svi_params* p) {
    if (!p) return SVI_ERR_INVALID;
    std::memset(p, 0, sizeof(*p));
    p->quality_level = 0.01;
    p->min_distance = 7.0;
    p->harris_k = 0.04;
    p->min_disparity_px = 0.01;
    p->max_corners = 1000;
    p->keypoint_size = 7.0f;
    p->search_range_px = 60.0f;
    p->match_cutoff = 100.0f;
    p->cutoff_stage1 = 25.0f;
    p->cutoff_stage2 = 50.0f;
    p->cutoff_stage3 = 50.0f;
    p->cutoff_original = 100.0f;
    p->max_candidates = 16384;
    p->chunk_frames = 0;
    p->max_queries = 16384;
    p->detector = SVI_DETECTOR_GFTT_HARRIS;
    p->fast_threshold = 10;
    p->fast_nonmax = 1;
    return SVI_SUCCESS;
}

const char* svi_status_text(int status) {
    switch (status) {
        case SVI_OK: return "ok";
        case SVI_TRI_RANGE: return "<CTriangulator>(getPointTriangulatedInRIGHT) insufficient search range";
        case SVI_TRI_NO_DESC: return "<CTriangulator>(getPointTriangulatedInRIGHT) could not compute descriptors";
        case SVI_TRI_NO_MATCH: return "<CTriangulator>(getPointTriangulatedInRIGHT) no match found";
        case SVI_TRI_DISTANCE: return "<CTriangulator>(getPointTriangulatedInRIGHT) matching distance";
        case SVI_TRI_ZERO_DISP: return "<CTriangulator>(getPointInLEFT) zero disparity";
        case SVI_TRI_BAD_ROI: return "search region outside image";
        case SVI_TRK_DEPTH: return "invalid depth";
        case SVI_TRK_STAGE1_DIST: return "insufficient matching distance";
        case SVI_TRK_TRI_DESC: return "triangulation descriptor mismatch";
        case SVI_TRK_OUT_OF_FOV: return "projection outside the field of view";
        case SVI_TRK_NO_FEATURES: return "no features detected";
        case SVI_TRK_NO_MATCHES: return "no matches found";
        case SVI_TRK_DESC: return "descriptor mismatch";
        case SVI_TRK_RANGE: return "out of tracking range";
        case SVI_EPI_OUT_OF_SIGHT: return "<CFundamentalMatcher>(getVisibleLandmarksFundamental) projection out of sight";
        case SVI_EPI_VERTICAL: return "<CFundamentalMatcher>(getVisibleLandmarksFundamental) vertical out of sight";
        case SVI_EPI_NEG_SLOPE: return "<CFundamentalMatcher>(getVisibleLandmarksFundamental) caught bad projection negative slope";
        case SVI_EPI_POS_SLOPE: return "<CFundamentalMatcher>(getVisibleLandmarksFundamental) caught bad projection positive slope";
        case SVI_EPI_ZERO_LEN: return "<CFundamentalMatcher>(getVisibleLandmarksFundamental) zero line length";
        case SVI_EPI_POOL_EMPTY: return "could not find a matching descriptor (empty key point pool)";
        case SVI_EPI_NO_MATCHES: return "could not find any matches (empty matches pool)";
        case SVI_EPI_DIST: return "could not find a matching descriptor";
        case SVI_EPI_ORIG_DIST: return "could not find a matching descriptor (ORIGINAL matching distance too big)";
        case SVI_EPI_NO_TRANSLATION: return "no translation since detection: epipolar search skipped";
        default: return "unknown status";
    }
}

const char* svi_brief_table_info(void) { return SVI_BRIEF_PATTERN_SOURCE; }

int svi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* svi_last_error(const svi_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void svi_destroy(svi_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < kMaxLanes; ++i) {
        Lane& l = ctx->lanes[i];
        if (l.stream) cudaStreamSynchronize(l.stream);
        void* ptrs[] = {l.box_l, l.box_r, l.box_ls, l.box_rs, l.frame_max, l.cand_count, l.cand, l.det_xy, l.kp_xy, l.n_det, l.n_kp,
                        l.g_head, l.g_next, l.g_state, l.img_l, l.img_r, l.mask, l.out.uv_l, l.out.uv_r, l.out.xyz,
                        l.out.desc_l, l.out.desc_r, l.out.dist, l.out.idx, l.out.status, l.bin_start, l.bin_slot, l.bin_kp};
        for (void* p : ptrs) if (p) cudaFree(p);
        for (cudaEvent_t e : l.ev) cudaEventDestroy(e);
        if (l.done) cudaEventDestroy(l.done);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    if (ctx->h_overflow) cudaFreeHost(const_cast<int*>(ctx->h_overflow));
    if (ctx->side_done) cudaEventDestroy(ctx->side_done);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->opt_dev) cudaFree(ctx->opt_dev);
    if (ctx->opt_pin) cudaFreeHost(ctx->opt_pin);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->small_block) cudaFree(ctx->small_block);
    if (ctx->pin_arena) cudaFreeHost(ctx->pin_arena);
    if (ctx->trk_img) cudaFree(ctx->trk_img);
    if (ctx->s3_items) cudaFree(ctx->s3_items);
    if (ctx->resp_one) cudaFree(ctx->resp_one);
    {
        void* rp[] = {ctx->roi.max, ctx->roi.cand, ctx->roi.det, ctx->roi.kp, ctx->roi.n_det, ctx->roi.n_kp, ctx->roi.rois, ctx->roi.s2};
        for (void* q : rp) if (q) cudaFree(q);
    }
    if (ctx->fork) cudaEventDestroy(ctx->fork);
    delete ctx;
}

int svi_create(const svi_camera* left, const svi_camera* right, const svi_params* params, int device, svi_ctx** out) {
    svi_ctx* ctx = nullptr;
    if (!left || !right || !out) return fail(nullptr, SVI_ERR_INVALID, "svi_create: null argument");
    *out = nullptr;
    int ndev = svi_device_count();
    if (ndev <= 0) return fail(nullptr, SVI_ERR_NO_DEVICE, "svi_create: no CUDA device (libsvi_gpu has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, SVI_ERR_INVALID, "svi_create: bad device index");
    if (left->width != right->width || left->height != right->height)
        return fail(nullptr, SVI_ERR_INVALID, "svi_create: left/right image sizes differ");
    if (left->width < 64 || left->height < 64 || left->width > 65535 || left->height > 65535)
        return fail(nullptr, SVI_ERR_INVALID, "svi_create: image size must be within [64, 65535]");
    svi_params p;
    if (params) p = *params; else svi_params_default(&p);
    if (p.max_corners <= 0 || p.max_corners > 65535) return fail(nullptr, SVI_ERR_INVALID, "svi_create: max_corners must be in [1, 65535]");
    if (p.max_candidates < 1024) p.max_candidates = 1024;
    if (p.max_queries < 1) p.max_queries = 1;
    {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return fail(nullptr, SVI_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    }
    ctx = new svi_ctx();
#undef CK
#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            svi_destroy(ctx);                                                                    \
            return fail(nullptr, SVI_ERR_CUDA, m_);                                              \
        }                                                                                        \
    } while (0)
    ctx->device = device;
    ctx->cam_l = *left;
    ctx->cam_r = *right;
    ctx->p = p;
    {
        int n_sm = 0;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && n_sm > 0) ctx->n_sm = n_sm;
    }
    ctx->W = (int)left->width;
    ctx->H = (int)left->height;
    ctx->dev_pitch = ctx->W;  // dense rows: staging copies are 1-D DMA transfers at full PCIe rate
    ctx->resp_pitch = align_up(ctx->W, 32);
    ctx->box_pitch = align_up(ctx->W, 64);
    const char* env_chunk = std::getenv("SVI_CHUNK_FRAMES");
    const char* env_lanes = std::getenv("SVI_LANES");
    // default: 64 KITTI-size frames per launch (measured sweep, profiles/r1_lane_sweep.txt), fewer for larger images so
    // that the scratch of a lane stays near half a gigabyte
    const int auto_chunk = (int)std::max<long long>(4, std::min<long long>(64, 30000000LL / ((long long)ctx->W * ctx->H)));
    ctx->chunk = p.chunk_frames > 0 ? p.chunk_frames : (env_chunk ? std::atoi(env_chunk) : auto_chunk);
    ctx->chunk = std::max(1, std::min(ctx->chunk, 4096));
    ctx->n_lanes = env_lanes ? std::max(1, std::min(std::atoi(env_lanes), kMaxLanes)) : 6;
    ctx->profiling = std::getenv("SVI_PROFILE") != nullptr;
    if (const char* e = std::getenv("SVI_MATCH_SPLIT")) {
        const int v = std::atoi(e);
        if (v == 1 || v == 2) ctx->match_split = v;
    }
    if (const char* e = std::getenv("SVI_MATCH_PRE")) ctx->match_pre = std::atoi(e) != 0;
    {   // The binned matcher's tile is laid out for scan lines of one pass (pool size ceil(range) + 1 <= 62 candidates) whose
        // ROI border 4 * size is a whole number of pixels: the reference's 60 px / size 7.  It pays on dense frames only
        // (measured, frames/s device-resident: C2, ~13 key-points per bin, 109.0 k vs 103.9 k for the per-key-point
        // kernels; C4, ~7 per bin, 132.2 k vs 136.0 k): expected key-points per bin = ~0.8 of max_corners (BRIEF's border
        // filter) spread over the image minus that border.  Anything else keeps the per-key-point kernels.
        int mode = 1;
        if (const char* e = std::getenv("SVI_MATCH_BINNED")) mode = std::atoi(e);
        const float b4 = 4.f * p.keypoint_size;
        const bool geom_ok = p.search_range_px > 0.f && std::ceil(p.search_range_px) + 1.f <= (float)BT_REACH && b4 == std::floor(b4) && b4 >= 1.f;
        const double per_bin = 0.8 * p.max_corners * (double)(BinWide::BIN_W * BinWide::BIN_H) /
                               std::max(1.0, (double)(ctx->W - 2 * kBriefBorder) * (double)(ctx->H - 2 * kBriefBorder));
        ctx->match_binned = geom_ok && (mode == 2 || (mode == 1 && per_bin >= 10.0));
        ctx->nbx = (ctx->W + BinWide::BIN_W - 1) / BinWide::BIN_W;
        ctx->n_bins = ctx->nbx * ((ctx->H + BinWide::BIN_H - 1) / BinWide::BIN_H);
        if (ctx->n_bins * (int)sizeof(uint32_t) > 40000) ctx->match_binned = false;   // the bin histogram must fit the default dynamic shared memory
    }
    ctx->trace = std::getenv("SVI_TRACE") != nullptr;
    int cap = 1024;
    while (cap < p.max_candidates) cap <<= 1;
    ctx->cand_cap = cap;
    ctx->raw_cap = cap * kRawFactor;

    // cornerHarris scale: 1 / ((1 << (ksize-1)) * blockSize) / 255 for 8-bit input (ksize 3, block 7)
    const double scale = 1.0 / (4.0 * 7.0 * 255.0);
    ctx->f1 = (float)scale;
    ctx->f0 = (float)(2.0 * scale);
    ctx->kf = (float)p.harris_k;

    // goodFeaturesToTrack's bucket grid; cell >= minDistance keeps the 3x3 neighbourhood sufficient
    SelectParams& sp = ctx->sel;
    sp.W = ctx->W; sp.H = ctx->H;
    sp.cand_cap = cap;
    sp.raw_cap = cap * kRawFactor;
    sp.quality = p.quality_level;
    sp.max_corners = p.max_corners;
    sp.filter = p.min_distance >= 1.0 ? 1 : 0;
    sp.cap_is_error = 0;
    if (p.detector == SVI_DETECTOR_FAST_9_16) {   // cv::FAST: no distance filter, no maxCorners cut
        sp.filter = 0;
        sp.cap_is_error = 1;
    } else if (p.detector != SVI_DETECTOR_GFTT_HARRIS) {
        delete ctx;
        return fail(nullptr, SVI_ERR_INVALID, "svi_create: unknown detector");
    }
    sp.min_dist_sq = p.min_distance * p.min_distance;
    int cell = std::max(1, (int)std::ceil(p.min_distance));
    while (((ctx->W + cell - 1) / cell) * ((ctx->H + cell - 1) / cell) > SEL_SMEM_CELLS && cap <= SEL_SMEM_KEYS && cell < 64) ++cell;
    sp.cell = cell;
    sp.cell_magic = cell < 256 ? (uint32_t)(((1u << 24) + cell - 1) / cell) : 0u;
    sp.min_dist_sq_ceil = (int)std::min(std::ceil(sp.min_dist_sq), 2.0e9);
    sp.gw = (ctx->W + cell - 1) / cell;
    sp.gh = (ctx->H + cell - 1) / cell;
    ctx->select_smem = (cap <= SEL_SMEM_KEYS) && (sp.gw * sp.gh <= SEL_SMEM_CELLS);

    // CTriangulator constants (src/core/CTriangulator.cpp:13-21)
    TriConst& tc = ctx->tc;
    const double f = left->P[0];
    tc.f_inv = 1.0 / f;
    tc.pu = left->P[2];
    tc.pv = left->P[6];
    tc.du_r_flipped = -right->P[3];
    tc.min_disp = p.min_disparity_px;
    tc.depth_min = tc.du_r_flipped / (double)left->width;
    tc.depth_max = tc.du_r_flipped / p.min_disparity_px;
    tc.width_left = (float)left->width;
    tc.width_right = (float)right->width;
    tc.match_cutoff = p.match_cutoff;

    CK(cudaFuncSetAttribute(harris_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HarrisSmem)));
    CK(cudaFuncSetAttribute(select_corners_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * SEL_SMEM_KEYS));
    CK(cudaFuncSetAttribute(stereo_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MATCH_SMEM));
    CK(cudaFuncSetAttribute(stereo_match_split_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MatchSplit<1>::SMEM));
    CK(cudaFuncSetAttribute(stereo_match_split_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MatchSplit<2>::SMEM));
    CK(cudaFuncSetAttribute(stereo_match_split_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MatchSplit<2>::SMEM));
    CK(cudaFuncSetAttribute(describe_left_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM));
    CK(cudaFuncSetAttribute(stereo_match_binned_kernel<BinWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, BinWide::M_SMEM));
    CK(cudaFuncSetAttribute(describe_left_binned_kernel<BinWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, BinWide::D_SMEM));
    CK(cudaFuncSetAttribute(triangulate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MATCH_SMEM));
    CK(cudaFuncSetAttribute(triangulate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MATCH_SMEM));
    CK(cudaFuncSetAttribute(track_stage1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MATCH_SMEM));
    CK(cudaFuncSetAttribute(track_stage2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MATCH_SMEM));
    CK(cudaFuncSetAttribute(track_stage2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MATCH_SMEM));
    CK(cudaFuncSetAttribute(track_stage3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MATCH_SMEM));
    // one shared-memory carve-out for every kernel of the pipeline: back-to-back kernels with different
    // carve-outs make the SMs drain and reconfigure between launches
    {
        const void* kernels[] = {(const void*)harris_box_kernel, (const void*)boxsum9_kernel,
                                 (const void*)select_corners_kernel<true>, (const void*)select_corners_kernel<false>,
                                 (const void*)stereo_match_kernel, (const void*)stereo_match_split_kernel<1, true>,
                                 (const void*)stereo_match_split_kernel<2, true>, (const void*)stereo_match_split_kernel<2, false>,
                                 (const void*)describe_left_kernel, (const void*)stereo_match_binned_kernel<BinWide>,
                                 (const void*)describe_left_binned_kernel<BinWide>, (const void*)bin_keypoints_kernel,
                                 (const void*)triangulate_kernel<true>,
                                 (const void*)triangulate_kernel<false>, (const void*)track_stage1_kernel,
                                 (const void*)track_stage2_kernel<true>, (const void*)track_stage2_kernel<false>,
                                 (const void*)track_stage3_kernel, (const void*)fast_candidates_kernel,
                                 (const void*)describe_kernel, (const void*)hamming_match_kernel};
        for (const void* k : kernels)
            CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }

    const size_t C = (size_t)ctx->chunk, HH = (size_t)ctx->H, MC = (size_t)p.max_corners;
    for (int i = 0; i < ctx->n_lanes; ++i) {
        Lane& l = ctx->lanes[i];
        CK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        CK(dmalloc(&l.box_l, C * HH * ctx->box_pitch));
        CK(dmalloc(&l.box_r, C * HH * ctx->box_pitch));
        CK(dmalloc(&l.box_ls, C * HH * ctx->box_pitch));
        CK(dmalloc(&l.box_rs, C * HH * ctx->box_pitch));
        CK(cudaMemset(l.box_ls, 0, sizeof(uint16_t) * C * HH * ctx->box_pitch));
        CK(cudaMemset(l.box_rs, 0, sizeof(uint16_t) * C * HH * ctx->box_pitch));
        {
            std::string merr;
            if (!make_box_map(&l.map_l, l.box_l, ctx->W, (int)(C * HH), ctx->box_pitch, &merr) ||
                !make_box_map(&l.map_ls, l.box_ls, ctx->W, (int)(C * HH), ctx->box_pitch, &merr) ||
                !make_box_map(&l.map_rs, l.box_rs, ctx->W, (int)(C * HH), ctx->box_pitch, &merr) ||
                !make_box_map(&l.map_r, l.box_r, ctx->W, (int)(C * HH), ctx->box_pitch, &merr) ||
                !make_box_map(&l.map_l_patch, l.box_l, ctx->W, (int)(C * HH), ctx->box_pitch, &merr, DL_COLS)) {
                svi_destroy(ctx);
                return fail(nullptr, SVI_ERR_CUDA, merr);
            }
        }
        if (ctx->match_binned) {
            std::string merr;
            CK(dmalloc(&l.bin_start, C * (size_t)(ctx->n_bins + 1)));
            CK(dmalloc(&l.bin_slot, C * MC));
            CK(dmalloc(&l.bin_kp, C * MC));
            if (setup_binned<BinWide>(ctx, l, C * HH, &merr) != SVI_SUCCESS) {
                svi_destroy(ctx);
                return fail(nullptr, SVI_ERR_CUDA, merr);
            }
        }
        CK(dmalloc(&l.frame_max, C));
        CK(dmalloc(&l.cand_count, C));
        CK(dmalloc(&l.cand, C * (size_t)ctx->raw_cap));
        CK(dmalloc(&l.det_xy, C * MC));
        CK(dmalloc(&l.kp_xy, C * MC));
        CK(dmalloc(&l.n_det, C));
        CK(dmalloc(&l.n_kp, C));
        if (!ctx->select_smem) {
            CK(dmalloc(&l.g_head, C * sp.gw * sp.gh));
            CK(dmalloc(&l.g_next, C * cap));
            CK(dmalloc(&l.g_state, C * cap));
        }
        CK(dmalloc(&l.img_l, C * HH * ctx->dev_pitch));
        CK(dmalloc(&l.img_r, C * HH * ctx->dev_pitch));
        CK(dmalloc(&l.mask, C * HH * ctx->dev_pitch));
        l.out.cap = p.max_corners;
        CK(dmalloc(&l.out.uv_l, C * MC * 2));
        CK(dmalloc(&l.out.uv_r, C * MC * 2));
        CK(dmalloc(&l.out.xyz, C * MC * 3));
        CK(dmalloc(&l.out.desc_l, C * MC * 32));
        CK(dmalloc(&l.out.desc_r, C * MC * 32));
        CK(dmalloc(&l.out.dist, C * MC));
        CK(dmalloc(&l.out.idx, C * MC));
        CK(dmalloc(&l.out.status, C * MC));
    }
    {
        int* h = nullptr;
        CK(cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(int), cudaHostAllocMapped));
        *h = 0;
        ctx->h_overflow = h;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->d_overflow), h, 0));
    }
    CK(cudaEventCreateWithFlags(&ctx->fork, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->side_done, cudaEventDisableTiming));
    CK(dmalloc(&ctx->trk_img, 2 * HH * ctx->dev_pitch));
    ctx->arena_bytes = (size_t)p.max_queries * 512 + (size_t)p.max_corners * 64 + 3 * HH * ctx->dev_pitch + (1 << 20);
    CK(cudaMalloc(reinterpret_cast<void**>(&ctx->arena), ctx->arena_bytes));
    if (ctx->arena_bytes <= (64u << 20)) CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->pin_arena), ctx->arena_bytes));
    // room for kSmallFrames frames: three image planes in, every output array out
    ctx->pin_bytes = (size_t)kSmallFrames * (3 * HH * ctx->dev_pitch + (size_t)p.max_corners * kOutBytesPerSlot + 64) + 16;
    if (ctx->pin_bytes <= (64u << 20)) CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->pin), ctx->pin_bytes));
    else ctx->pin_bytes = 0;
    {   // layout (descending alignment): xyz | uv_l | uv_r | dist | idx | n_kp | n_det | desc_l | desc_r | status
        const size_t S = (size_t)kSmallFrames * MC;
        ctx->small_bytes = S * kOutBytesPerSlot + 2 * sizeof(int) * kSmallFrames;
        CK(cudaMalloc(reinterpret_cast<void**>(&ctx->small_block), ctx->small_bytes));
        unsigned char* q = ctx->small_block;
        StereoOutDev& o = ctx->small_out;
        o.cap = p.max_corners;
        o.xyz = reinterpret_cast<double*>(q); q += S * 24;
        o.uv_l = reinterpret_cast<float*>(q); q += S * 8;
        o.uv_r = reinterpret_cast<float*>(q); q += S * 8;
        o.dist = reinterpret_cast<int*>(q); q += S * 4;
        o.idx = reinterpret_cast<int*>(q); q += S * 4;
        ctx->small_n_kp = reinterpret_cast<int*>(q); q += sizeof(int) * kSmallFrames;
        ctx->small_n_det = reinterpret_cast<int*>(q); q += sizeof(int) * kSmallFrames;
        o.desc_l = q; q += S * 32;
        o.desc_r = q; q += S * 32;
        o.status = q;
    }
    *out = ctx;
    return SVI_SUCCESS;
#undef CK
#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(ctx, SVI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)
}

int svi_stereo_frames_device(svi_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t pitch, size_t frame_stride,
                             int n_frames, const uint8_t* masks, const svi_stereo_result* out, void* cuda_stream) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!left || !right || !out || n_frames < 0 || (int)pitch < ctx->W || frame_stride < pitch * (size_t)ctx->H)
        return fail(ctx, SVI_ERR_INVALID, "svi_stereo_frames_device: bad argument");
    if (out->capacity_per_frame < ctx->p.max_corners) return fail(ctx, SVI_ERR_CAPACITY, "svi_stereo_frames_device: capacity_per_frame < max_corners");
    if (!out->n_keypoints || !out->uv_left || !out->uv_right || !out->xyz_left || !out->desc_left || !out->desc_right ||
        !out->distance || !out->match_index || !out->status)
        return fail(ctx, SVI_ERR_INVALID, "svi_stereo_frames_device: null output array");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t user = reinterpret_cast<cudaStream_t>(cuda_stream);
    CK(cudaEventRecord(ctx->fork, user));
    for (int i = 0; i < ctx->n_lanes; ++i) CK(cudaStreamWaitEvent(ctx->lanes[i].stream, ctx->fork, 0));
    const FrameGeom g = make_geom(ctx, (int)pitch, frame_stride);
    StereoOutDev o;
    o.cap = out->capacity_per_frame;
    o.uv_l = out->uv_left; o.uv_r = out->uv_right; o.xyz = out->xyz_left;
    o.desc_l = out->desc_left; o.desc_r = out->desc_right;
    o.dist = out->distance; o.idx = out->match_index; o.status = out->status;
    int chunk_id = 0, rc = SVI_SUCCESS;
    for (int f0 = 0; f0 < n_frames; f0 += ctx->chunk, ++chunk_id) {
        Lane& l = ctx->lanes[ctx->serial ? 0 : chunk_id % ctx->n_lanes];
        const int nf = std::min(ctx->chunk, n_frames - f0);
        int* n_det = out->n_detected ? out->n_detected + f0 : l.n_det;
        rc = run_pipeline(ctx, l, left + (size_t)f0 * frame_stride, right + (size_t)f0 * frame_stride,
                          masks ? masks + (size_t)f0 * frame_stride : nullptr, g, nf, o, f0, out->n_keypoints + f0, n_det);
        if (rc != SVI_SUCCESS) break;
    }
    // join the lanes back into the caller's stream -- also after a failure, so that nothing queued so far can outlive
    // the caller's next use of its buffers
    for (int i = 0; i < ctx->n_lanes; ++i) {
        if (cudaEventRecord(ctx->lanes[i].done, ctx->lanes[i].stream) == cudaSuccess)
            cudaStreamWaitEvent(user, ctx->lanes[i].done, 0);
    }
    return rc;
}

int svi_stereo_frames(svi_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t pitch, size_t frame_stride,
                      int n_frames, const uint8_t* masks, svi_stereo_result* out) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!left || !right || !out || n_frames < 0 || (int)pitch < ctx->W || (n_frames > 1 && frame_stride < pitch * (size_t)ctx->H))
        return fail(ctx, SVI_ERR_INVALID, "svi_stereo_frames: bad argument");
    if (out->capacity_per_frame < ctx->p.max_corners) return fail(ctx, SVI_ERR_CAPACITY, "svi_stereo_frames: capacity_per_frame < max_corners");
    if (!out->n_keypoints || !out->uv_left || !out->uv_right || !out->xyz_left || !out->desc_left || !out->desc_right ||
        !out->distance || !out->match_index || !out->status)
        return fail(ctx, SVI_ERR_INVALID, "svi_stereo_frames: null output array");
    CK(cudaSetDevice(ctx->device));
    const int W = ctx->W, H = ctx->H, MC = ctx->p.max_corners, cap = out->capacity_per_frame;
    const size_t dstride = (size_t)H * ctx->dev_pitch;
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, dstride);
    const bool dense = (frame_stride == pitch * (size_t)H);
    if (n_frames > 0 && n_frames <= std::min(kSmallFrames, ctx->chunk) && ctx->pin)
        return small_call(ctx, left, right, pitch, frame_stride, n_frames, masks, nullptr, 0, out);
    int chunk_id = 0;
    for (int f0 = 0; f0 < n_frames; f0 += ctx->chunk, ++chunk_id) {
        Lane& l = ctx->lanes[chunk_id % ctx->n_lanes];
        cudaStream_t s = l.stream;
        const int nf = std::min(ctx->chunk, n_frames - f0);
        const uint8_t* srcs[3] = {left, right, masks};
        uint8_t* dsts[3] = {l.img_l, l.img_r, l.mask};
        for (int k = 0; k < 3; ++k) {
            if (!srcs[k]) continue;
            if (dense && (int)pitch == ctx->dev_pitch) {
                CK(cudaMemcpyAsync(dsts[k], srcs[k] + (size_t)f0 * frame_stride, (size_t)nf * frame_stride, cudaMemcpyHostToDevice, s));
            } else if (dense) {
                CK(cudaMemcpy2DAsync(dsts[k], ctx->dev_pitch, srcs[k] + (size_t)f0 * frame_stride, pitch, W, (size_t)H * nf,
                                     cudaMemcpyHostToDevice, s));
            } else {
                for (int f = 0; f < nf; ++f)
                    CK(cudaMemcpy2DAsync(dsts[k] + f * dstride, ctx->dev_pitch, srcs[k] + (size_t)(f0 + f) * frame_stride,
                                         pitch, W, H, cudaMemcpyHostToDevice, s));
            }
        }
        int rc = run_pipeline(ctx, l, l.img_l, l.img_r, masks ? l.mask : nullptr, g, nf, l.out, 0, l.n_kp, l.n_det);
        if (rc != SVI_SUCCESS) return rc;
        const size_t o0 = (size_t)f0 * cap;
#define D2H(dst, src, elem)                                                                                             \
    do {                                                                                                                \
        if (cap == MC)                                                                                                  \
            CK(cudaMemcpyAsync((dst) + o0 * (elem), (src), (size_t)nf * MC * (elem) * sizeof(*(dst)),                   \
                               cudaMemcpyDeviceToHost, s));                                                             \
        else                                                                                                            \
            CK(cudaMemcpy2DAsync((dst) + o0 * (elem), (size_t)cap * (elem) * sizeof(*(dst)), (src),                     \
                                 (size_t)MC * (elem) * sizeof(*(dst)), (size_t)MC * (elem) * sizeof(*(dst)), nf,        \
                                 cudaMemcpyDeviceToHost, s));                                                           \
    } while (0)
        D2H(out->uv_left, l.out.uv_l, 2);
        D2H(out->uv_right, l.out.uv_r, 2);
        D2H(out->xyz_left, l.out.xyz, 3);
        D2H(out->desc_left, l.out.desc_l, 32);
        D2H(out->desc_right, l.out.desc_r, 32);
        D2H(out->distance, l.out.dist, 1);
        D2H(out->match_index, l.out.idx, 1);
        D2H(out->status, l.out.status, 1);
#undef D2H
        CK(cudaMemcpyAsync(out->n_keypoints + f0, l.n_kp, sizeof(int) * nf, cudaMemcpyDeviceToHost, s));
        if (out->n_detected) CK(cudaMemcpyAsync(out->n_detected + f0, l.n_det, sizeof(int) * nf, cudaMemcpyDeviceToHost, s));
    }
    int rc = sync_lanes(ctx);
    if (rc != SVI_SUCCESS) return rc;
    return check_overflow(ctx);
}

int svi_mask_active_landmarks(svi_ctx* ctx, const float* centres_xy, int n_centres, uint8_t* mask, size_t pitch) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!mask || n_centres < 0 || (n_centres > 0 && !centres_xy) || (int)pitch < ctx->W)
        return fail(ctx, SVI_ERR_INVALID, "svi_mask_active_landmarks: bad argument");
    if (n_centres > ctx->p.max_queries) return fail(ctx, SVI_ERR_CAPACITY, "svi_mask_active_landmarks: n_centres > max_queries");
    CK(cudaSetDevice(ctx->device));
    Lane& l = ctx->lanes[0];
    int rc = build_mask(ctx, l, centres_xy, n_centres);
    if (rc != SVI_SUCCESS) return rc;
    CK(cudaMemcpy2DAsync(mask, pitch, l.mask, ctx->dev_pitch, ctx->W, ctx->H, cudaMemcpyDeviceToHost, l.stream));
    CK(cudaStreamSynchronize(l.stream));
    return SVI_SUCCESS;
}

int svi_stereo_frame_masked(svi_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t pitch, const float* centres_xy,
                            int n_centres, svi_stereo_result* out) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!left || !right || !out || n_centres < 0 || (n_centres > 0 && !centres_xy) || (int)pitch < ctx->W)
        return fail(ctx, SVI_ERR_INVALID, "svi_stereo_frame_masked: bad argument");
    if (out->capacity_per_frame < ctx->p.max_corners) return fail(ctx, SVI_ERR_CAPACITY, "svi_stereo_frame_masked: capacity_per_frame < max_corners");
    if (!out->n_keypoints || !out->uv_left || !out->uv_right || !out->xyz_left || !out->desc_left || !out->desc_right ||
        !out->distance || !out->match_index || !out->status)
        return fail(ctx, SVI_ERR_INVALID, "svi_stereo_frame_masked: null output array");
    if (n_centres > ctx->p.max_queries) return fail(ctx, SVI_ERR_CAPACITY, "svi_stereo_frame_masked: n_centres > max_queries");
    if (!ctx->pin) return fail(ctx, SVI_ERR_UNSUPPORTED, "svi_stereo_frame_masked: frame too large for the pinned bounce buffer");
    CK(cudaSetDevice(ctx->device));
    static const float kNone[2] = {-1.0e9f, -1.0e9f};   // no landmark yet: an all-255 mask
    return small_call(ctx, left, right, pitch, pitch * (size_t)ctx->H, 1, nullptr, n_centres > 0 ? centres_xy : kNone,
                      n_centres > 0 ? n_centres : 1, out);
}

int svi_harris_response(svi_ctx* ctx, const uint8_t* img, size_t pitch, float* response) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!img || !response || (int)pitch < ctx->W) return fail(ctx, SVI_ERR_INVALID, "svi_harris_response: bad argument");
    CK(cudaSetDevice(ctx->device));
    Lane& l = ctx->lanes[0];
    cudaStream_t s = l.stream;
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, (size_t)ctx->H * ctx->dev_pitch);
    CK(cudaMemcpy2DAsync(l.img_l, ctx->dev_pitch, img, pitch, ctx->W, ctx->H, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(l.frame_max, 0, sizeof(uint32_t), s));
    if (!ctx->resp_one) CK(dmalloc(&ctx->resp_one, (size_t)ctx->H * ctx->resp_pitch));
    harris_box_kernel<<<harris_grid(g.W, g.H, 1), HT_THREADS, sizeof(HarrisSmem), s>>>(
        l.img_l, nullptr, g, ctx->f1, ctx->f0, ctx->kf, ctx->p.quality_level, ctx->resp_one, nullptr, nullptr, l.frame_max, nullptr,
        nullptr, 0, nullptr, nullptr, g.H);
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(response, sizeof(float) * ctx->W, ctx->resp_one, sizeof(float) * ctx->resp_pitch, sizeof(float) * ctx->W,
                         ctx->H, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SVI_SUCCESS;
}

int svi_detect(svi_ctx* ctx, const uint8_t* img, size_t pitch, size_t frame_stride, int n_frames, const uint8_t* masks,
               float* xy, int32_t* counts) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!img || !xy || !counts || n_frames < 0 || (int)pitch < ctx->W) return fail(ctx, SVI_ERR_INVALID, "svi_detect: bad argument");
    CK(cudaSetDevice(ctx->device));
    Lane& l = ctx->lanes[0];
    cudaStream_t s = l.stream;
    const int W = ctx->W, H = ctx->H, MC = ctx->p.max_corners;
    const size_t dstride = (size_t)H * ctx->dev_pitch;
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, dstride);
    std::vector<ushort2> h_xy((size_t)ctx->chunk * MC);
    for (int f0 = 0; f0 < n_frames; f0 += ctx->chunk) {
        const int nf = std::min(ctx->chunk, n_frames - f0);
        for (int f = 0; f < nf; ++f) {
            CK(cudaMemcpy2DAsync(l.img_l + f * dstride, ctx->dev_pitch, img + (size_t)(f0 + f) * frame_stride, pitch, W, H,
                                 cudaMemcpyHostToDevice, s));
            if (masks)
                CK(cudaMemcpy2DAsync(l.mask + f * dstride, ctx->dev_pitch, masks + (size_t)(f0 + f) * frame_stride, pitch, W, H,
                                     cudaMemcpyHostToDevice, s));
        }
        const uint8_t* d_mask = masks ? l.mask : nullptr;
        CK(cudaMemsetAsync(l.frame_max, 0, sizeof(uint32_t) * nf, s));
        CK(cudaMemsetAsync(l.cand_count, 0, sizeof(int) * nf, s));
        const dim3 tiles((W + HT_W - 1) / HT_W, (H + HT_H - 1) / HT_H, nf);
        const bool fast = ctx->p.detector == SVI_DETECTOR_FAST_9_16;
        if (fast) {
            fast_candidates_kernel<<<tiles, HT_THREADS, 0, s>>>(l.img_l, d_mask, g, ctx->p.fast_threshold, ctx->p.fast_nonmax, l.cand,
                                                                l.cand_count, ctx->raw_cap);
        } else {
            harris_box_kernel<<<harris_grid(W, H, nf), HT_THREADS, sizeof(HarrisSmem), s>>>(
                l.img_l, d_mask, g, ctx->f1, ctx->f0, ctx->kf, ctx->p.quality_level, nullptr, nullptr, nullptr, l.frame_max, l.cand,
                l.cand_count, ctx->raw_cap, nullptr, nullptr, g.H);
        }
        const uint32_t* fmax = fast ? nullptr : l.frame_max;
        if (ctx->select_smem)
            select_corners_kernel<true><<<nf, SEL_THREADS, 13 * SEL_SMEM_KEYS, s>>>(l.cand, l.cand_count, fmax, ctx->sel, nullptr, nullptr,
                                                                                  nullptr, l.det_xy, l.n_det, l.kp_xy, l.n_kp,
                                                                                  ctx->d_overflow, nullptr);
        else
            select_corners_kernel<false><<<nf, SEL_THREADS, 4 * SEL_SMEM_CELLS, s>>>(l.cand, l.cand_count, fmax, ctx->sel, l.g_head, l.g_next, l.g_state,
                                                                   l.det_xy, l.n_det, l.kp_xy, l.n_kp, ctx->d_overflow, nullptr);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_xy.data(), l.det_xy, sizeof(ushort2) * (size_t)nf * MC, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(counts + f0, l.n_det, sizeof(int) * nf, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int f = 0; f < nf; ++f)
            for (int i = 0; i < counts[f0 + f]; ++i) {
                xy[((size_t)(f0 + f) * MC + i) * 2] = (float)h_xy[(size_t)f * MC + i].x;
                xy[((size_t)(f0 + f) * MC + i) * 2 + 1] = (float)h_xy[(size_t)f * MC + i].y;
            }
    }
    return check_overflow(ctx);
}

int svi_describe(svi_ctx* ctx, const uint8_t* img, size_t pitch, const float* xy, int n, uint8_t* desc32, uint8_t* kept) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!img || !xy || !desc32 || !kept || n < 0 || (int)pitch < ctx->W) return fail(ctx, SVI_ERR_INVALID, "svi_describe: bad argument");
    if (n > ctx->p.max_queries) return fail(ctx, SVI_ERR_CAPACITY, "svi_describe: n > max_queries");
    if (n == 0) return SVI_SUCCESS;
    CK(cudaSetDevice(ctx->device));
    Lane& l = ctx->lanes[0];
    cudaStream_t s = l.stream;
    ctx->arena_used = 0; ctx->pending.clear();
    int rc = stage_box(ctx, img, pitch, l.img_l, l.box_l, l.box_ls, s);
    if (rc != SVI_SUCCESS) return rc;
    float* d_xy; uint8_t* d_desc; uint8_t* d_kept;
    UP(d_xy, xy, (size_t)n * 2);
    UP(d_desc, (const uint8_t*)nullptr, (size_t)n * 32);
    UP(d_kept, (const uint8_t*)nullptr, (size_t)n);
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, (size_t)ctx->H * ctx->dev_pitch);
    describe_kernel<<<(n + 3) / 4, 128, 0, s>>>(l.box_l, g, d_xy, n, d_desc, d_kept);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(desc32, d_desc, (size_t)n * 32, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(kept, d_kept, (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SVI_SUCCESS;
}

int svi_match_hamming(svi_ctx* ctx, const uint8_t* query32, int n_query, const uint8_t* train32, int n_train, int32_t* index,
                      int32_t* distance) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!query32 || !index || !distance || n_query < 0 || n_train < 0 || (n_train > 0 && !train32))
        return fail(ctx, SVI_ERR_INVALID, "svi_match_hamming: bad argument");
    if ((size_t)(n_query + n_train) * 40 > ctx->arena_bytes) return fail(ctx, SVI_ERR_CAPACITY, "svi_match_hamming: raise max_queries");
    if (n_query == 0) return SVI_SUCCESS;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->lanes[0].stream;
    ctx->arena_used = 0; ctx->pending.clear();
    uint8_t* d_q; uint8_t* d_t; int* d_i; int* d_d;
    UP(d_q, query32, (size_t)n_query * 32);
    UP(d_t, n_train > 0 ? train32 : (const uint8_t*)nullptr, (size_t)std::max(n_train, 1) * 32);
    UP(d_i, (const int*)nullptr, (size_t)n_query);
    UP(d_d, (const int*)nullptr, (size_t)n_query);
    hamming_match_kernel<<<(n_query + 3) / 4, 128, 0, s>>>(d_q, n_query, d_t, n_train, d_i, d_d);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(index, d_i, sizeof(int) * n_query, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(distance, d_d, sizeof(int) * n_query, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SVI_SUCCESS;
}

int svi_match_epipolar(svi_ctx* ctx, const uint8_t* query32, const float* query_xy, int n_query, const uint8_t* train32,
                       const float* train_xy, int n_train, float band_v, float min_disparity, float max_disparity,
                       int32_t* index, int32_t* distance, int32_t* second_distance) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!query32 || !query_xy || !index || !distance || !second_distance || n_query < 0 || n_train < 0 ||
        (n_train > 0 && (!train32 || !train_xy)))
        return fail(ctx, SVI_ERR_INVALID, "svi_match_epipolar: bad argument");
    if ((size_t)(n_query + n_train) * 48 + (size_t)n_query * 16 > ctx->arena_bytes)
        return fail(ctx, SVI_ERR_CAPACITY, "svi_match_epipolar: raise max_queries");
    if (n_query == 0) return SVI_SUCCESS;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->lanes[0].stream;
    ctx->arena_used = 0; ctx->pending.clear();
    uint8_t* d_q; uint8_t* d_t; float* d_qxy; float* d_txy; int* d_i; int* d_d; int* d_s;
    UP(d_q, query32, (size_t)n_query * 32);
    UP(d_qxy, query_xy, (size_t)n_query * 2);
    UP(d_t, n_train > 0 ? train32 : (const uint8_t*)nullptr, (size_t)std::max(n_train, 1) * 32);
    UP(d_txy, n_train > 0 ? train_xy : (const float*)nullptr, (size_t)std::max(n_train, 1) * 2);
    UP(d_i, (const int*)nullptr, (size_t)n_query);
    UP(d_d, (const int*)nullptr, (size_t)n_query);
    UP(d_s, (const int*)nullptr, (size_t)n_query);
    epipolar_match_kernel<<<(n_query + 3) / 4, 128, 0, s>>>(d_q, d_qxy, n_query, d_t, d_txy, n_train, band_v, min_disparity,
                                                            max_disparity, d_i, d_d, d_s);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(index, d_i, sizeof(int) * n_query, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(distance, d_d, sizeof(int) * n_query, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(second_distance, d_s, sizeof(int) * n_query, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SVI_SUCCESS;
}

int svi_triangulate_right(svi_ctx* ctx, const uint8_t* img_right, size_t pitch, int n, const float* top_left,
                          const float* uv_left, const uint8_t* desc_left, float keypoint_size, svi_tri_result* out) {
    if (!ctx) return SVI_ERR_INVALID;
    return triangulate_common(ctx, false, img_right, pitch, n, nullptr, top_left, uv_left, desc_left, keypoint_size, out);
}

int svi_triangulate_left(svi_ctx* ctx, const uint8_t* img_left, size_t pitch, int n, const float* search_range,
                         const float* top_left, const float* uv_right, const uint8_t* desc_right, float keypoint_size,
                         svi_tri_result* out) {
    if (!ctx) return SVI_ERR_INVALID;
    return triangulate_common(ctx, true, img_left, pitch, n, search_range, top_left, uv_right, desc_right, keypoint_size, out);
}

int svi_point_in_left(svi_ctx* ctx, int n, const float* uv_left, const float* uv_right, double* xyz_left, uint8_t* status) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!uv_left || !uv_right || !xyz_left || !status || n < 0) return fail(ctx, SVI_ERR_INVALID, "svi_point_in_left: bad argument");
    if (n > ctx->p.max_queries) return fail(ctx, SVI_ERR_CAPACITY, "svi_point_in_left: n > max_queries");
    if (n == 0) return SVI_SUCCESS;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->lanes[0].stream;
    ctx->arena_used = 0; ctx->pending.clear();
    float* d_l; float* d_r; double* d_xyz; uint8_t* d_st;
    UP(d_l, uv_left, (size_t)n * 2);
    UP(d_r, uv_right, (size_t)n * 2);
    UP(d_xyz, (const double*)nullptr, (size_t)n * 3);
    UP(d_st, (const uint8_t*)nullptr, (size_t)n);
    point_in_left_kernel<<<(n + 127) / 128, 128, 0, s>>>(ctx->tc, n, d_l, d_r, d_xyz, d_st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(xyz_left, d_xyz, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(status, d_st, (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SVI_SUCCESS;
}

int svi_track_landmarks(svi_ctx* ctx, const uint8_t* img_left, const uint8_t* img_right, size_t pitch,
                        const double* T_world_to_left, const svi_landmarks* lm, int n, double motion_scaling,
                        svi_track_result* out) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!lm) return fail(ctx, SVI_ERR_INVALID, "svi_track_landmarks: bad argument");
    const bool any3 = lm->uv_reference_left || lm->desc_reference_left || lm->T_left_to_world_at_detection;
    return svi_track_landmarks_stages(ctx, img_left, img_right, pitch, T_world_to_left, lm, n, motion_scaling,
                                      SVI_STAGE_1 | SVI_STAGE_2 | (any3 ? SVI_STAGE_3 : 0), out);
}

int svi_track_landmarks_stages(svi_ctx* ctx, const uint8_t* img_left, const uint8_t* img_right, size_t pitch,
                               const double* T_world_to_left, const svi_landmarks* lm, int n, double motion_scaling,
                               uint32_t stage_mask, svi_track_result* out) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!img_left || !img_right || !T_world_to_left || !lm || !out || n < 0 || (int)pitch < ctx->W ||
        !(stage_mask & (SVI_STAGE_1 | SVI_STAGE_2 | SVI_STAGE_3)) || (stage_mask & ~(uint32_t)(SVI_STAGE_1 | SVI_STAGE_2 | SVI_STAGE_3)))
        return fail(ctx, SVI_ERR_INVALID, "svi_track_landmarks: bad argument");
    if (!lm->xyz_world || !lm->last_desc_left || !lm->last_desc_right || !lm->last_disparity || !lm->keypoint_size ||
        !out->status || !out->stage || !out->uv_left || !out->uv_right || !out->xyz_left || !out->desc_left || !out->desc_right)
        return fail(ctx, SVI_ERR_INVALID, "svi_track_landmarks: null array");
    if (n > ctx->p.max_queries) return fail(ctx, SVI_ERR_CAPACITY, "svi_track_landmarks: n > max_queries");
    // the reference computes the scaling as min(1 + ..., 5) (CTrackerGT.cpp:157) and asserts it positive; the search
    // windows grow with it, so an absurd value is rejected here instead of producing image-sized windows
    if (!(motion_scaling >= 0.0 && motion_scaling <= 16.0))
        return fail(ctx, SVI_ERR_INVALID, "svi_track_landmarks: motion_scaling must be within [0, 16]");
    const bool stage3 = (stage_mask & SVI_STAGE_3) != 0;
    if (stage3 && !(lm->uv_reference_left && lm->desc_reference_left && lm->T_left_to_world_at_detection))
        return fail(ctx, SVI_ERR_INVALID, "svi_track_landmarks: stage 3 needs all three reference arrays");
    if (n == 0) return SVI_SUCCESS;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_track_scratch(ctx);
    if (rc != SVI_SUCCESS) return rc;
    Lane& l = ctx->lanes[0];
    cudaStream_t s = l.stream;
    ctx->arena_used = 0; ctx->pending.clear();
    using clk = std::chrono::steady_clock;
    clk::time_point tp[6];
    auto tick = [&](int i) { if (ctx->trace) tp[i] = clk::now(); };
    tick(0);
    const size_t plane = (size_t)ctx->H * ctx->dev_pitch;
    // both images as planes 0 / 1 of one buffer (the window-mode detector indexes them by plane)
    // (RIGHT goes up and gets its box sums on the side stream, beside LEFT's)
    CK(cudaEventRecord(ctx->fork, s));
    CK(cudaStreamWaitEvent(ctx->side, ctx->fork, 0));
    rc = stage_box(ctx, img_right, pitch, ctx->trk_img + plane, l.box_r, l.box_rs, ctx->side, 1);
    if (rc == SVI_SUCCESS) rc = stage_box(ctx, img_left, pitch, ctx->trk_img, l.box_l, l.box_ls, s);
    if (cudaEventRecord(ctx->side_done, ctx->side) != cudaSuccess || cudaStreamWaitEvent(s, ctx->side_done, 0) != cudaSuccess)
        rc = rc == SVI_SUCCESS ? fail(ctx, SVI_ERR_CUDA, "svi_track_landmarks: joining the side stream failed") : rc;
    if (rc != SVI_SUCCESS) { cudaStreamSynchronize(ctx->side); cudaStreamSynchronize(s); return rc; }
    // every input array goes through the pinned mirror of the arena and up in ONE copy; the outputs form one contiguous
    // range behind them: one memset, one copy back
    double* d_xyzw; uint8_t* d_dl; uint8_t* d_dr; float* d_disp; float* d_size;
    ctx->batch_uploads = ctx->pin_arena != nullptr;
    struct Unbatch { svi_ctx* c; ~Unbatch() { c->batch_uploads = false; } } unbatch{ctx};
    UP(d_xyzw, lm->xyz_world, (size_t)n * 3);
    UP(d_dl, lm->last_desc_left, (size_t)n * 32);
    UP(d_dr, lm->last_desc_right, (size_t)n * 32);
    UP(d_disp, lm->last_disparity, (size_t)n);
    UP(d_size, lm->keypoint_size, (size_t)n);
    Stage3Extra ex{nullptr, nullptr};
    uint8_t* d_orig = nullptr;
    if (stage3) {
        double* d_uvref; double* d_tdet;
        UP(d_uvref, lm->uv_reference_left, (size_t)n * 2);
        UP(d_tdet, lm->T_left_to_world_at_detection, (size_t)n * 16);
        UP(d_orig, lm->desc_reference_left, (size_t)n * 32);
        ex.uv_ref = d_uvref; ex.T_det = d_tdet;
    }
    tick(1);
    if (ctx->batch_uploads) CK(cudaMemcpyAsync(ctx->arena, ctx->pin_arena, ctx->arena_used, cudaMemcpyHostToDevice, s));
    ctx->batch_uploads = false;
    const size_t out_begin = (ctx->arena_used + 255) & ~size_t(255);
    TrackOutDev o;
    UP(o.status, (const uint8_t*)nullptr, (size_t)n);
    UP(o.stage, (const uint8_t*)nullptr, (size_t)n);
    UP(o.uv_l, (const float*)nullptr, (size_t)n * 2);
    UP(o.uv_r, (const float*)nullptr, (size_t)n * 2);
    UP(o.xyz, (const double*)nullptr, (size_t)n * 3);
    UP(o.desc_l, (const uint8_t*)nullptr, (size_t)n * 32);
    UP(o.desc_r, (const uint8_t*)nullptr, (size_t)n * 32);
    const size_t out_end = ctx->arena_used;
    CK(cudaMemsetAsync(ctx->arena + out_begin, 0, out_end - out_begin, s));
    TrackConst k;
    for (int i = 0; i < 12; ++i) { k.T[i] = T_world_to_left[i]; k.PL[i] = ctx->cam_l.P[i]; k.PR[i] = ctx->cam_r.P[i]; }
    k.tri_scale = (float)(1.0 + motion_scaling);
    k.cutoff1 = ctx->p.cutoff_stage1;
    k.stage1_match = (stage_mask & SVI_STAGE_1) ? 1 : 0;
    LandmarksDev ld{d_xyzw, d_dl, d_dr, d_disp, d_size};
    const FrameGeom g = make_geom(ctx, ctx->dev_pitch, plane);
    // The whole cascade is ONE stream of kernels: every stage reads the stage / status bytes the previous one left on
    // the device and plans its own work items there; the host only waits once, for the results.
    if (stage_mask & (SVI_STAGE_1 | SVI_STAGE_2)) {
        // ---- stage 1 LEFT / RIGHT for every landmark (or only its field-of-view gate when stage 2 runs alone)
        track_stage1_kernel<<<(n + MATCH_WARPS - 1) / MATCH_WARPS, MATCH_WARPS * 32, MATCH_SMEM, s>>>(l.box_l, l.box_r, l.map_l, l.map_ls, l.map_r, l.map_rs, g,
                                                                                                      ctx->tc, k, ld, n, o);
        CK(cudaGetLastError());
    } else {
        // stage 3 alone (trackEpipolar :828-1020): no gate, every landmark is an untracked candidate
        CK(cudaMemsetAsync(o.status, SVI_TRK_STAGE1_DIST, (size_t)n, s));
        CK(cudaMemsetAsync(o.stage, 0, (size_t)n, s));
    }
    // ---- stage 2 LEFT, then stage 2 RIGHT, for what is still untracked
    for (int side = 0; side < 2 && (stage_mask & SVI_STAGE_2); ++side) {
        rc = track_stage2_side(ctx, l, g, n, T_world_to_left, motion_scaling, side == 0, ld, o);
        if (rc != SVI_SUCCESS) { cudaStreamSynchronize(s); return rc; }
    }
    // ---- stage 3 (epipolar line in LEFT)
    if (stage3) {
        rc = track_stage3_all(ctx, l, g, n, T_world_to_left, motion_scaling, ld, ex, d_orig, o);
        if (rc != SVI_SUCCESS) { cudaStreamSynchronize(s); return rc; }
    }
    tick(2);
    int rd = SVI_SUCCESS;
    if (ctx->pin_arena) {   // one copy for the whole output range, scattered to the caller's arrays after the synchronisation
        cudaError_t e = cudaMemcpyAsync(ctx->pin_arena + out_begin, ctx->arena + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) rd = fail(ctx, SVI_ERR_CUDA, std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e));
        auto mirror = [&](const void* d) { return ctx->pin_arena + (static_cast<const unsigned char*>(d) - ctx->arena); };
        ctx->pending.push_back({out->status, mirror(o.status), (size_t)n});
        ctx->pending.push_back({out->stage, mirror(o.stage), (size_t)n});
        ctx->pending.push_back({out->uv_left, mirror(o.uv_l), sizeof(float) * 2 * n});
        ctx->pending.push_back({out->uv_right, mirror(o.uv_r), sizeof(float) * 2 * n});
        ctx->pending.push_back({out->xyz_left, mirror(o.xyz), sizeof(double) * 3 * n});
        ctx->pending.push_back({out->desc_left, mirror(o.desc_l), (size_t)32 * n});
        ctx->pending.push_back({out->desc_right, mirror(o.desc_r), (size_t)32 * n});
    } else {
        rd = download(ctx, out->status, o.status, (size_t)n, s);
        if (rd == SVI_SUCCESS) rd = download(ctx, out->stage, o.stage, (size_t)n, s);
        if (rd == SVI_SUCCESS) rd = download(ctx, out->uv_left, o.uv_l, sizeof(float) * 2 * n, s);
        if (rd == SVI_SUCCESS) rd = download(ctx, out->uv_right, o.uv_r, sizeof(float) * 2 * n, s);
        if (rd == SVI_SUCCESS) rd = download(ctx, out->xyz_left, o.xyz, sizeof(double) * 3 * n, s);
        if (rd == SVI_SUCCESS) rd = download(ctx, out->desc_left, o.desc_l, (size_t)32 * n, s);
        if (rd == SVI_SUCCESS) rd = download(ctx, out->desc_right, o.desc_r, (size_t)32 * n, s);
    }
    if (ctx->trace) { cudaStreamSynchronize(s); tick(3); }
    if (rd == SVI_SUCCESS) rd = flush_downloads(ctx, s);
    else cudaStreamSynchronize(s);
    tick(4);
    if (rd != SVI_SUCCESS) return rd;
    rd = check_overflow(ctx);
    tick(5);
    if (ctx->trace) {
        auto us = [&](int a, int b) { return std::chrono::duration<double, std::micro>(tp[b] - tp[a]).count(); };
        std::fprintf(stderr, "svi_track_landmarks n=%d: stage images + mirror inputs %.0f us | enqueue cascade %.0f us | wait %.0f us | "
                             "scatter outputs %.0f us | overflow check %.0f us | total %.0f us\n",
                     n, us(0, 1), us(1, 2), us(2, 3), us(3, 4), us(4, 5), us(0, 5));
    }
    return rd;
}

int svi_optimize_landmarks(svi_ctx* ctx, const svi_landmark_measurements* in, int n, svi_optimize_result* out) {
    if (!ctx) return SVI_ERR_INVALID;
    if (!in || !out || n < 0 || !out->xyz_world || !out->outcome || !out->average_squared_error)
        return fail(ctx, SVI_ERR_INVALID, "svi_optimize_landmarks: bad argument");
    if (n == 0) return SVI_SUCCESS;
    if (!in->xyz_world_guess || !in->first || in->n_poses < 0) return fail(ctx, SVI_ERR_INVALID, "svi_optimize_landmarks: null array");
    const int m = in->first[n];
    if (in->first[0] != 0 || m < 0) return fail(ctx, SVI_ERR_INVALID, "svi_optimize_landmarks: first[] must start at 0 and end at the measurement count");
    for (int i = 0; i < n; ++i)
        if (in->first[i + 1] < in->first[i]) return fail(ctx, SVI_ERR_INVALID, "svi_optimize_landmarks: first[] must not decrease");
    if (m > 0 && (!in->pose_index || !in->uv_left || !in->uv_right || !in->proj_world_to_left || !in->proj_world_to_right || in->n_poses == 0))
        return fail(ctx, SVI_ERR_INVALID, "svi_optimize_landmarks: null measurement array");
    for (int k = 0; k < m; ++k)
        if (in->pose_index[k] < 0 || in->pose_index[k] >= in->n_poses) return fail(ctx, SVI_ERR_INVALID, "svi_optimize_landmarks: pose_index out of range");
    CK(cudaSetDevice(ctx->device));
    using clk = std::chrono::steady_clock;
    const clk::time_point t_begin = clk::now();
    // layout: [xyz_guess | proj_left | proj_right | first | pose_index | uv_left | uv_right] -> [xyz | avg | iterations | outcome]
    const size_t P = (size_t)in->n_poses;
    size_t off = 0;
    auto take = [&off](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    const size_t o_guess = take(sizeof(double) * 3 * n), o_pl = take(sizeof(double) * 12 * P), o_pr = take(sizeof(double) * 12 * P);
    const size_t o_first = take(sizeof(int) * ((size_t)n + 1)), o_pose = take(sizeof(int) * (size_t)m);
    const size_t o_uvl = take(sizeof(float) * 2 * (size_t)m), o_uvr = take(sizeof(float) * 2 * (size_t)m);
    const size_t in_bytes = off;
    const size_t o_xyz = take(sizeof(double) * 3 * n), o_avg = take(sizeof(double) * n), o_it = take(sizeof(int) * (size_t)n), o_out = take((size_t)n);
    const size_t total = off;
    if (total > ctx->opt_bytes) {   // grows to twice the need: a tracker's measurement lists get longer frame by frame
        if (ctx->opt_dev) cudaFree(ctx->opt_dev);
        if (ctx->opt_pin) cudaFreeHost(ctx->opt_pin);
        ctx->opt_dev = nullptr; ctx->opt_pin = nullptr; ctx->opt_bytes = 0;
        const size_t want = std::max<size_t>(2 * total, 1u << 20);
        CK(cudaMalloc(reinterpret_cast<void**>(&ctx->opt_dev), want));
        CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->opt_pin), want));
        ctx->opt_bytes = want;
    }
    unsigned char* hp = ctx->opt_pin;
    std::memcpy(hp + o_guess, in->xyz_world_guess, sizeof(double) * 3 * n);
    std::memcpy(hp + o_first, in->first, sizeof(int) * ((size_t)n + 1));
    if (m > 0) {
        std::memcpy(hp + o_pl, in->proj_world_to_left, sizeof(double) * 12 * P);
        std::memcpy(hp + o_pr, in->proj_world_to_right, sizeof(double) * 12 * P);
        std::memcpy(hp + o_pose, in->pose_index, sizeof(int) * (size_t)m);
        std::memcpy(hp + o_uvl, in->uv_left, sizeof(float) * 2 * (size_t)m);
        std::memcpy(hp + o_uvr, in->uv_right, sizeof(float) * 2 * (size_t)m);
    }
    cudaStream_t s = ctx->lanes[0].stream;
    const clk::time_point t_staged = clk::now();
    CK(cudaMemcpyAsync(ctx->opt_dev, hp, in_bytes, cudaMemcpyHostToDevice, s));
    unsigned char* d = ctx->opt_dev;
    LandmarkOptIn li{reinterpret_cast<const double*>(d + o_guess), reinterpret_cast<const int*>(d + o_first), reinterpret_cast<const int*>(d + o_pose),
                     reinterpret_cast<const float*>(d + o_uvl), reinterpret_cast<const float*>(d + o_uvr), reinterpret_cast<const double*>(d + o_pl),
                     reinterpret_cast<const double*>(d + o_pr), in->n_poses};
    LandmarkOptOut lo{reinterpret_cast<double*>(d + o_xyz), d + o_out, reinterpret_cast<double*>(d + o_avg), reinterpret_cast<int*>(d + o_it)};
    optimize_landmarks_kernel<<<(n + kOptWarps - 1) / kOptWarps, kOptWarps * 32, 0, s>>>(li, n, lo);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hp + o_xyz, d + o_xyz, total - o_xyz, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::memcpy(out->xyz_world, hp + o_xyz, sizeof(double) * 3 * n);
    std::memcpy(out->average_squared_error, hp + o_avg, sizeof(double) * n);
    std::memcpy(out->outcome, hp + o_out, (size_t)n);
    if (out->iterations) std::memcpy(out->iterations, hp + o_it, sizeof(int) * (size_t)n);
    if (const char* dump = std::getenv("SVI_OPT_DUMP")) {   // diagnostics: the landmark with the most iterations, in facade_demo --landmark format
        const int* it = reinterpret_cast<const int*>(hp + o_it);
        int worst = 0;
        for (int i = 1; i < n; ++i) if (it[i] > it[worst]) worst = i;
        if (std::FILE* f = std::fopen(dump, "w")) {
            std::fprintf(f, "%.17g %.17g %.17g\n", in->xyz_world_guess[3 * worst], in->xyz_world_guess[3 * worst + 1], in->xyz_world_guess[3 * worst + 2]);
            for (int k = in->first[worst]; k < in->first[worst + 1]; ++k) {
                for (int q = 0; q < 12; ++q) std::fprintf(f, "%.17g ", in->proj_world_to_left[12 * (size_t)in->pose_index[k] + q]);
                for (int q = 0; q < 12; ++q) std::fprintf(f, "%.17g ", in->proj_world_to_right[12 * (size_t)in->pose_index[k] + q]);
                std::fprintf(f, "%.9g %.9g %.9g %.9g\n", in->uv_left[2 * k], in->uv_left[2 * k + 1], in->uv_right[2 * k], in->uv_right[2 * k + 1]);
            }
            std::fprintf(f, "# iterations %d outcome %d\n", it[worst], (int)hp[o_out + worst]);
            std::fclose(f);
        }
    }
    if (ctx->trace) {
        const clk::time_point t_end = clk::now();
        auto us = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        std::fprintf(stderr, "svi_optimize_landmarks n=%d m=%d poses=%d: check + stage inputs %.0f us | copy up, kernel, copy back %.0f us | total %.0f us\n",
                     n, m, in->n_poses, us(t_begin, t_staged), us(t_staged, t_end), us(t_begin, t_end));
    }
    return check_overflow(ctx);
}

int svi_set_profiling(svi_ctx* ctx, int enable) {
    if (!ctx) return SVI_ERR_INVALID;
    for (int i = 0; i < ctx->n_lanes; ++i) cudaStreamSynchronize(ctx->lanes[i].stream);
    collect_timings(ctx);
    for (int s = 0; s < kStages; ++s) { ctx->stage_ms[s] = 0.0; ctx->stage_launches[s] = 0; }
    ctx->profiling = enable != 0;
    ctx->serial = enable == 2;
    if (ctx->profiling)
        for (int i = 0; i < ctx->n_lanes; ++i) reserve_events(ctx->lanes[i], ctx->serial && i == 0 ? 8 * kEventPool : kEventPool);
    return SVI_SUCCESS;
}

int svi_check_overflow(svi_ctx* ctx) {
    if (!ctx) return SVI_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < ctx->n_lanes; ++i) CK(cudaStreamSynchronize(ctx->lanes[i].stream));
    return check_overflow(ctx);
}

int svi_stage_timings(svi_ctx* ctx, const char** names, double* total_ms, int64_t* launches, int capacity) {
    if (!ctx || !names || !total_ms || !launches) return SVI_ERR_INVALID;
    for (int i = 0; i < ctx->n_lanes; ++i) cudaStreamSynchronize(ctx->lanes[i].stream);
    collect_timings(ctx);
    int n = std::min(capacity, kStages);
    for (int i = 0; i < n; ++i) {
        names[i] = kStageNames[i];
        total_ms[i] = ctx->stage_ms[i];
        launches[i] = ctx->stage_launches[i];
    }
    return n;
}

// ---- frame-partitioned batches over several GPUs: stereo pairs are independent, so a batch of F
// frames is cut into contiguous ranges [g*F/G, (g+1)*F/G), one per device; each range runs through its own svi_ctx on
// its own host thread and writes a disjoint slice of the caller's arrays.  No collective, no peer access: the result is
// byte-for-byte the one a single device produces.
struct svi_multi {
    std::vector<svi_ctx*> ctx;
    std::vector<int> devices;
    std::string err;
};

int svi_multi_create(const svi_camera* left, const svi_camera* right, const svi_params* params, const int* devices, int n_devices,
                     svi_multi** out) {
    if (!left || !right || !out || n_devices < 1 || n_devices > 64) return fail(nullptr, SVI_ERR_INVALID, "svi_multi_create: bad argument");
    *out = nullptr;
    svi_multi* m = new svi_multi();
    for (int g = 0; g < n_devices; ++g) {
        svi_ctx* c = nullptr;
        const int dev = devices ? devices[g] : g;
        const int rc = svi_create(left, right, params, dev, &c);
        if (rc != SVI_SUCCESS) {   // g_create_error holds the reason
            for (svi_ctx* q : m->ctx) svi_destroy(q);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
        m->devices.push_back(dev);
    }
    *out = m;
    return SVI_SUCCESS;
}

void svi_multi_destroy(svi_multi* m) {
    if (!m) return;
    for (svi_ctx* c : m->ctx) svi_destroy(c);
    delete m;
}

const char* svi_multi_last_error(const svi_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }
int svi_multi_device_count(const svi_multi* m) { return m ? (int)m->ctx.size() : 0; }

int svi_multi_frame_range(const svi_multi* m, int n_frames, int part, int* first, int* count) {
    if (!m || !first || !count || n_frames < 0 || part < 0 || part >= (int)m->ctx.size()) return SVI_ERR_INVALID;
    const long long G = (long long)m->ctx.size();
    const long long f0 = (long long)part * n_frames / G, f1 = (long long)(part + 1) * n_frames / G;
    *first = (int)f0;
    *count = (int)(f1 - f0);
    return SVI_SUCCESS;
}

int svi_multi_stereo_frames(svi_multi* m, const uint8_t* left, const uint8_t* right, size_t pitch, size_t frame_stride, int n_frames,
                            const uint8_t* masks, svi_stereo_result* out) {
    if (!m) return SVI_ERR_INVALID;
    if (!left || !right || !out || n_frames < 0) { m->err = "svi_multi_stereo_frames: bad argument"; return SVI_ERR_INVALID; }
    const int G = (int)m->ctx.size();
    std::vector<int> rc(G, SVI_SUCCESS);
    std::vector<std::thread> workers;
    const size_t cap = (size_t)out->capacity_per_frame;
    for (int g = 0; g < G; ++g) {
        int f0 = 0, nf = 0;
        svi_multi_frame_range(m, n_frames, g, &f0, &nf);
        if (nf == 0) continue;
        workers.emplace_back([=, &rc]() {
            svi_stereo_result sub = *out;   // the same arrays, advanced to this range's first frame
            const size_t o0 = (size_t)f0 * cap;
            sub.n_keypoints = out->n_keypoints ? out->n_keypoints + f0 : nullptr;
            sub.n_detected = out->n_detected ? out->n_detected + f0 : nullptr;
            sub.uv_left = out->uv_left ? out->uv_left + o0 * 2 : nullptr;
            sub.uv_right = out->uv_right ? out->uv_right + o0 * 2 : nullptr;
            sub.xyz_left = out->xyz_left ? out->xyz_left + o0 * 3 : nullptr;
            sub.desc_left = out->desc_left ? out->desc_left + o0 * 32 : nullptr;
            sub.desc_right = out->desc_right ? out->desc_right + o0 * 32 : nullptr;
            sub.distance = out->distance ? out->distance + o0 : nullptr;
            sub.match_index = out->match_index ? out->match_index + o0 : nullptr;
            sub.status = out->status ? out->status + o0 : nullptr;
            rc[g] = svi_stereo_frames(m->ctx[g], left + (size_t)f0 * frame_stride, right + (size_t)f0 * frame_stride, pitch, frame_stride, nf,
                                      masks ? masks + (size_t)f0 * frame_stride : nullptr, &sub);
        });
    }
    for (std::thread& t : workers) t.join();
    for (int g = 0; g < G; ++g)
        if (rc[g] != SVI_SUCCESS) {
            m->err = "device " + std::to_string(m->devices[g]) + ": " + svi_last_error(m->ctx[g]);
            return rc[g];
        }
    return SVI_SUCCESS;
}

int svi_kernels_per_chunk(const svi_ctx* ctx, int n_frames) {
    if (!ctx || n_frames <= 0) return SVI_ERR_INVALID;
    const long long slots = (long long)std::min(n_frames, ctx->chunk) * ctx->p.max_corners;
    const bool pre = ctx->match_pre && slots / (2LL * ctx->n_sm * 9) >= MATCH_KP_PER_WARP;
    // detector, RIGHT box sums, corner selection, matcher (+ LEFT descriptors as a kernel of their own, + the bin sort)
    return 4 + (pre ? 1 : 0) + (pre && ctx->match_binned ? 1 : 0);
}

int svi_config(const svi_ctx* ctx, int32_t* chunk_frames, int32_t* n_lanes, int32_t* select_in_smem) {
    if (!ctx) return SVI_ERR_INVALID;
    if (chunk_frames) *chunk_frames = ctx->chunk;
    if (n_lanes) *n_lanes = ctx->n_lanes;
    if (select_in_smem) *select_in_smem = ctx->select_smem ? 1 : 0;
    return SVI_SUCCESS;
}

}  // extern "C"
