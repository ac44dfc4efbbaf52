// api.cu -- the C ABI of libmltcnn.so (include/mltcnn.h): context, weight blob, buffers, launch sequence.
// Replaces the inline libtorch/OpenCV block of EncCu::xCompressCU (EncCu.cpp:803-926).  No CPU fallback.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mlt_internal.h"

using namespace mlt;

namespace {

constexpr uint32_t MLTW_MAGIC = 0x57544C4Du;
enum : uint32_t {
    SEC_CONV1_F32 = 0x001, SEC_CONV1_UMMA = 0x002, SEC_STEM_CONV1 = 0x003, SEC_STEM5_W = 0x004, SEC_STEM5_CORR = 0x005, SEC_W_F16 = 0x100, SEC_BIAS_FUSED = 0x200, SEC_W_F32 = 0x300, SEC_BIAS = 0x400,
    SEC_SC_W_F16 = 0x500, SEC_SC_W_F32 = 0x600, SEC_SC_BIAS = 0x700, SEC_FC_W = 0x800, SEC_FC_B = 0x900, SEC_BIAS_MMA = 0xA00, SEC_X_W_F16 = 0xB00,
    SEC_W_SPLIT = 0xC00, SEC_X_SPLIT = 0xD00, // layer3 (4 ways) and layer2 (4 ways) weights packed for the channel-split kernels
    SEC_W_SPLIT8 = 0xE00, SEC_X_SPLIT8 = 0xF00 // layer3 split 8 ways
};

constexpr int NCONV = 16, NACT = 17, CTU = MLT_CTU_SIZE;
constexpr size_t CTU_IN_ELEMS = (size_t)2 * CTU * CTU; // org + pred planes
constexpr size_t CTU_PACKED_BYTES = MLT_CTU_PACKED10_BYTES; // the same as a 10-bit packed stream (pack10.cu)
static_assert(CTU_PACKED_BYTES == CTU_IN_ELEMS * 10 / 8, "packed CTU size");

// forward order of the 16 3x3 convs after conv1 (arch.py:247-254 with BasicBlock [2,2,2,2])
const LayerDesc kLayers[NCONV] = {
    {32, 32, 2, 64, -1},   {32, 32, 1, 64, 0},    {32, 32, 1, 64, -1},   {32, 32, 1, 64, -1},
    {32, 64, 2, 32, -1},   {64, 64, 1, 32, 1},    {64, 64, 1, 32, -1},   {64, 64, 1, 32, -1},
    {64, 128, 2, 16, -1},  {128, 128, 1, 16, 2},  {128, 128, 1, 16, -1}, {128, 128, 1, 16, -1},
    {128, 256, 2, 8, -1},  {256, 256, 1, 8, 3},   {256, 256, 1, 8, -1},  {256, 256, 1, 8, -1},
};

// activation a: 0 = conv1 out, 1 + li = output of conv li.  (H, C) of each.
inline void act_shape(int a, int &h, int &c)
{
    if (a == 0) { h = 128; c = 32; return; }
    h = kLayers[a - 1].hout;
    c = kLayers[a - 1].cout;
}
inline size_t act_elems(int a)
{
    int h, c;
    act_shape(a, h, c);
    return (size_t)h * h * c;
}
inline ActLayout act_layout(int a) // fp16 product-path layout of activation a (conv_umma.cuh)
{
    return a == 0 ? ActLayout{128, 32, 1, 0} : conv_umma_out_layout(a - 1);
}

struct Section { const uint8_t *dev = nullptr; size_t bytes = 0; };

} // namespace

struct mlt_ctx {
    int device = 0, max_batch = 0, engine = 0, num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    static constexpr int MAX_CHUNKS = 16;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}; // "chunk i of the input batch has landed in HBM"
    uint8_t *d_blob = nullptr;
    size_t blob_bytes = 0;
    Section sec[0x1000];
    // Two sets of activation buffers (+ the tensor maps over them): large batches are cut in two slices that run on two
    // streams, so the drain / launch / prologue gap between the 17 kernels of one slice is filled by kernels of the other.
    // Set 0 holds max_batch images (every single-slice call and all debug reads use it), set 1 half of that.
    struct ActSet {
        __half *act_h[NACT] = {}; // [0] (conv1's full output, parity-planar) exists only for the unfused engine / debug reads
        __half *act0q = nullptr;  // conv1's output at even rows / columns (input of layer0.0's shortcut), dense [n][4][64][64][8]
        ConvParams conv_p[NCONV]; // tensor maps + weight pointers of every tcgen05 conv, built once at create
        ConvParams conv_split[NCONV]; // layer3 (and layer2) again for the channel-split kernels (small batches): split-packed weights, 4 ways
        ConvParams conv_split8[NCONV]; // layer3 split 8 ways (tiny batches)
        float *gap_part[3] = {};  // pool partial sums written by convs 7 / 11 / 15: [cap][8 | 2 | 1 tiles][4][64 | 128 | 256]
        int cap = 0;              // images
    } set[2];
    cudaStream_t stream2 = nullptr;            // second compute stream (slice 1)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // mlt_predict_batch_device runs asynchronously on the CALLER's stream but uses the context's activation sets, d_ctus and
    // stream2: ev_dev marks the end of the last device call so that the next call -- on any stream, through any entry point --
    // is ordered after it instead of racing on those buffers
    cudaEvent_t ev_dev = nullptr;
    bool dev_issued = false, dev_pending_host = false;
    float *act_f[NACT] = {};
    float *scratch_f = nullptr; // shortcut-conv output of the fp32 engine
    // Host-batch slot: the device / pinned buffers one host batch lives in.  Slot 0 = the buffers below (every synchronous
    // call); slot 1 is allocated on the first mlt_submit_batch_dense so that two batches can be in flight (the H2D of
    // batch k + 1 runs while batch k computes).
    struct HostSlot {
        int16_t *d_in = nullptr;
        uint8_t *d_packed = nullptr; // 10-bit packed transport (mlt_*_packed10): H2D lands here, unpack10_kernel fills d_in; on first use
        CtuDev *d_ctus = nullptr, *h_ctus = nullptr;
        mlt_result *d_out = nullptr, *h_out = nullptr;
        cudaEvent_t done = nullptr;
        int n = 0;
        bool busy = false;
    } slot[2];
    uint64_t submitted = 0, collected = 0;
    int16_t *d_in = nullptr, *h_in = nullptr; // dense [max_batch][2][128][128]
    CtuDev *d_ctus = nullptr, *h_ctus = nullptr;
    mlt_result *d_out = nullptr, *h_out = nullptr;
    float *d_dbg = nullptr; // mlt_debug_stage / mlt_debug_activation staging
    size_t dbg_bytes = 0;
    int16_t *d_pic = nullptr; // picture original luma (mlt_begin_picture)
    size_t pic_capacity = 0;
    int pic_pitch = 0, pic_w = 0, pic_h = 0, pic_poc = 0;
    bool pic_valid = false;
    int16_t *d_ref = nullptr; // reference-picture luma of the frame-level pre-pass (mlt_predict_picture), same pitch as d_pic
    size_t ref_capacity = 0;
    PicCtu *d_pic_ctus = nullptr, *h_pic_ctus = nullptr; // eligible CTUs of the picture: position + integer MV
    bool ref_valid = false;   // d_ref holds a reference plane uploaded for the picture begun
    unsigned *d_me_cost = nullptr, *d_me_best = nullptr, *h_me_best = nullptr; // block matching: [n][(2R+1)^2] scratch, [n] winners
    int16_t *d_me_mv = nullptr, *h_me_mv = nullptr;
    int last_n = 0;
    uint64_t launches = 0;
    bool profiling = false;
    cudaEvent_t prof_ev[NCONV + 3] = {}; // boundaries of: stage+conv1, 16 convs, head
    cudaStream_t prof_stream = nullptr;
    bool prof_valid = false;
    std::string err;
};

namespace {

int fail(mlt_ctx *c, int rc, const char *fmt, ...)
{
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return rc;
}

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return fail(c, e_ == cudaErrorMemoryAllocation ? MLT_E_NOMEM : MLT_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int load_blob(mlt_ctx *c, const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return fail(c, MLT_E_IO, "cannot open weight blob '%s'", path);
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    if (sz < 0) { fclose(f); return fail(c, MLT_E_IO, "cannot determine the size of '%s'", path); }
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> raw((size_t)sz);
    const size_t got = raw.empty() ? 0 : fread(raw.data(), 1, raw.size(), f);
    fclose(f);
    if (got != raw.size() || raw.size() < 32) return fail(c, MLT_E_IO, "short read on '%s'", path);
    uint32_t hdr[4];
    uint64_t total;
    memcpy(hdr, raw.data(), 16);
    memcpy(&total, raw.data() + 16, 8);
    if (hdr[0] != MLTW_MAGIC || hdr[1] != 2 || hdr[2] != 128 || total != raw.size() || 32 + (size_t)hdr[3] * 24 > raw.size())
        return fail(c, MLT_E_FORMAT, "'%s' is not an MLTW v2 blob for the 128x128 CTU model", path);
    CU(cudaMalloc(&c->d_blob, raw.size()));
    c->blob_bytes = raw.size();
    CU(cudaMemcpy(c->d_blob, raw.data(), raw.size(), cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < hdr[3]; i++) {
        uint32_t id, dt;
        uint64_t off, nb;
        const uint8_t *e = raw.data() + 32 + (size_t)i * 24;
        memcpy(&id, e, 4); memcpy(&dt, e + 4, 4); memcpy(&off, e + 8, 8); memcpy(&nb, e + 16, 8);
        if (id >= 0x1000 || off > raw.size() || nb > raw.size() - off || (off & 255) || c->sec[id].dev != nullptr) return fail(c, MLT_E_FORMAT, "bad section table in '%s'", path);
        c->sec[id].dev = c->d_blob + off;
        c->sec[id].bytes = nb;
    }
    // every section this architecture needs must be present with the exact size
    auto need = [&](uint32_t id, size_t bytes) { return c->sec[id].dev != nullptr && c->sec[id].bytes == bytes; };
    bool ok = need(SEC_CONV1_F32, 9 * 2 * 32 * 4) && need(SEC_CONV1_UMMA, 2 * 4 * 32 * 16) && need(SEC_STEM_CONV1, 4 * 1024) &&
              need(SEC_STEM5_W, 7 * 1024) && need(SEC_STEM5_CORR, (2 * 5 * 2 * 32 + 2 * 32 + 32) * 4);
    for (int li = 0; li < NCONV && ok; li++) {
        const LayerDesc &L = kLayers[li];
        ok = need(SEC_W_F16 + li, (size_t)9 * L.cin * L.cout * 2) && need(SEC_W_F32 + li, (size_t)9 * L.cin * L.cout * 4) &&
             need(SEC_BIAS + li, (size_t)L.cout * 4) && need(SEC_BIAS_FUSED + li, (size_t)L.cout * 4) &&
             need(SEC_BIAS_MMA + li, (size_t)L.cout * 32);
        if (ok && li >= CONV_L2SPLIT_FIRST) ok = need(SEC_W_SPLIT + li, (size_t)9 * L.cin * L.cout * 2);
        if (ok && li >= CONV_SPLIT_FIRST) ok = need(SEC_W_SPLIT8 + li, (size_t)9 * L.cin * L.cout * 2);
        if (ok && (li & 1)) { // second conv of a block: extra operand = shortcut conv (first block) or identity
            const int xc = L.sc >= 0 ? kLayers[li - 1].cin : L.cout;
            ok = need(SEC_X_W_F16 + li, (size_t)xc * L.cout * 2 * (L.sc >= 0 ? 2 : 1)); // shortcut weights: hi + lo
            if (ok && li >= CONV_L2SPLIT_FIRST) ok = need(SEC_X_SPLIT + li, (size_t)xc * L.cout * 2 * (L.sc >= 0 ? 2 : 1));
            if (ok && li >= CONV_SPLIT_FIRST) ok = need(SEC_X_SPLIT8 + li, (size_t)xc * L.cout * 2 * (L.sc >= 0 ? 2 : 1));
        }
        if (ok && L.sc >= 0) {
            const int csc = kLayers[li - 1].cin;
            ok = need(SEC_SC_W_F32 + L.sc, (size_t)csc * L.cout * 4) && need(SEC_SC_BIAS + L.sc, (size_t)L.cout * 4);
        }
    }
    static const int fin[3] = {66, 130, 258}, fout[3] = {2, 3, 4};
    for (int i = 0; i < 3 && ok; i++) ok = need(SEC_FC_W + i, (size_t)fin[i] * fout[i] * 4) && need(SEC_FC_B + i, (size_t)fout[i] * 4);
    if (!ok) return fail(c, MLT_E_FORMAT, "'%s': missing or mis-sized section", path);
    return MLT_OK;
}

template <typename T>
const T *secp(const mlt_ctx *c, uint32_t id) { return reinterpret_cast<const T *>(c->sec[id].dev); }

int ensure_f32_buffers(mlt_ctx *c)
{
    if (c->act_f[0]) return MLT_OK;
    for (int a = 0; a < NACT; a++) CU(cudaMalloc(&c->act_f[a], act_elems(a) * c->max_batch * sizeof(float)));
    CU(cudaMalloc(&c->scratch_f, (size_t)64 * 64 * 32 * c->max_batch * sizeof(float)));
    return MLT_OK;
}

int ensure_dbg(mlt_ctx *c, size_t bytes)
{
    if (c->dbg_bytes >= bytes) return MLT_OK;
    if (c->d_dbg) cudaFree(c->d_dbg);
    c->d_dbg = nullptr;
    c->dbg_bytes = 0;
    CU(cudaMalloc(&c->d_dbg, bytes));
    c->dbg_bytes = bytes;
    return MLT_OK;
}

int ensure_act0(mlt_ctx *c, mlt_ctx::ActSet &S) // conv1's full output: only the unfused engine and debug reads materialise it
{
    if (S.act_h[0]) return MLT_OK;
    const ActLayout L = act_layout(0);
    const size_t bytes = L.unit_elems() * L.units_for(S.cap) * sizeof(__half);
    CU(cudaMalloc(&S.act_h[0], bytes));
    CU(cudaMemset(S.act_h[0], 0, bytes));
    ConvParams &p = S.conv_p[0];
    CU(conv_umma_prepare(0, &p, S.act_h[0], L, nullptr, nullptr, (size_t)S.cap));
    p.w = secp<__half>(c, SEC_W_F16 + 0);
    p.bias = secp<__half>(c, SEC_BIAS_MMA + 0);
    p.bias_f32 = secp<float>(c, SEC_BIAS_FUSED + 0);
    p.x_w = nullptr;
    p.out = S.act_h[1];
    p.relu = 1;
    return MLT_OK;
}

__global__ void dense_descs_kernel(CtuDev *ctus, const int16_t *orgpred, const int32_t *pocqp, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CtuDev d;
    d.org = orgpred + (size_t)i * CTU_IN_ELEMS;
    d.pred = d.org + (size_t)CTU * CTU;
    d.org_stride = d.pred_stride = CTU;
    d.poc = pocqp[2 * i];
    d.qp = pocqp[2 * i + 1];
    ctus[i] = d;
}

// The whole network on `n` CTUs described by device array `ctus`; results to device array `out`.
int run_network(mlt_ctx *c, const CtuDev *ctus, int n, mlt_result *out, cudaStream_t s, int set_idx = 0)
{
    mlt_ctx::ActSet &S = c->set[set_idx];
    if (n > S.cap) return fail(c, MLT_E_BATCH, "slice of %d CTUs exceeds activation set %d (%d)", n, set_idx, S.cap);
    HeadParams hp;
    hp.ctus = ctus; hp.out = out; hp.n = n;
    for (int i = 0; i < 3; i++) { hp.fc_w[i] = secp<float>(c, SEC_FC_W + i); hp.fc_b[i] = secp<float>(c, SEC_FC_B + i); }
    if (c->engine != 1) {
        // ---- product path: fused stem (staging + conv1 + layer0.0.conv1), 15 tcgen05 implicit-GEMM convs, head
        const bool prof = c->profiling && set_idx == 0;
        int ev = 0;
        if (prof) { CU(cudaEventRecord(c->prof_ev[ev++], s)); c->prof_stream = s; c->prof_valid = false; }
        int first = 0;
        if (c->engine == 0) {
            // fused stem: staging + conv1 + layer0.0.conv1; conv1's 1 MiB / CTU output never reaches HBM
            static const bool old_stem = getenv("MLT_STEM_OLD") != nullptr; // A/B switch: the round-1 stem (conv1 and layer0.0.conv1 as two MMA stages)
            if (old_stem)
                CU(launch_stem_umma(ctus, n, secp<__half>(c, SEC_STEM_CONV1), secp<__half>(c, SEC_W_F16 + 0), secp<float>(c, SEC_BIAS_FUSED + 0),
                                    S.act0q, S.act_h[1], c->num_sms, s));
            else
                CU(launch_stem5_umma(ctus, n, secp<__half>(c, SEC_STEM5_W), secp<float>(c, SEC_STEM5_CORR), secp<float>(c, SEC_STEM5_CORR) + 704,
                                     S.act0q, S.act_h[1], c->num_sms, s)); // bias: the section's last 32 floats (corrected for W5's fp16 rounding)
            c->launches++;
            if (prof) { CU(cudaEventRecord(c->prof_ev[ev++], s)); CU(cudaEventRecord(c->prof_ev[ev++], s)); }
            first = 1;
        } else {
            // unfused tcgen05 engine (cross-check of the stem): standalone conv1, its even/even quarter copied out
            int rc = ensure_act0(c, S);
            if (rc) return rc;
            CU(launch_conv1_umma(ctus, n, secp<__half>(c, SEC_CONV1_UMMA), S.act_h[0], s));
            c->launches++;
            const size_t q = (size_t)4 * 64 * 64 * 8 * sizeof(__half);
            CU(cudaMemcpy2DAsync(S.act0q, q, S.act_h[0], 4 * q, q, n, cudaMemcpyDeviceToDevice, s));
            if (prof) CU(cudaEventRecord(c->prof_ev[ev++], s));
        }
        // MLT_CHAIN=1 (experiment, off by default): for one or two CTUs run convs 1..15 as ONE cluster kernel -- cluster barriers instead
        // of 15 kernel boundaries.  Per-tile arithmetic is the per-layer kernels' own, so results are bit-identical (GPU test); measured
        // on B200 it is NOT faster (151 vs 139 us per call): PDL already overlaps each kernel's prologue and weight fetch with its
        // predecessor, the chain serialises them, and layer3 is bound by the ~27 GB/s one SM's bulk copies reach (profiles/r02/README.md)
        static const bool use_chain = getenv("MLT_CHAIN") != nullptr && getenv("MLT_NO_CHAIN") == nullptr;
        if (first == 1 && n <= CHAIN_MAX_IMAGES && !prof && use_chain && !getenv("MLT_TRACE_LAYER")) {
            ChainParams cp;
            for (int li = 1; li < NCONV; li++) {
                cp.p[li - 1] = li >= CONV_SPLIT_FIRST ? S.conv_split8[li] : (li >= CONV_L2SPLIT_FIRST ? S.conv_split[li] : S.conv_p[li]);
                cp.p[li - 1].nimg = n;
            }
            static const bool chain_trace = getenv("MLT_CHAIN_TRACE") != nullptr; // debug: per-layer %globaltimer stamps of CTA 0 -> stderr
            static unsigned long long *d_ct = nullptr;
            if (chain_trace && !d_ct) CU(cudaMalloc(&d_ct, 64 * sizeof(unsigned long long)));
            cp.trace = chain_trace ? d_ct : nullptr;
            static const int chain_trace_layer = getenv("MLT_CHAIN_TRACE_LAYER") ? atoi(getenv("MLT_CHAIN_TRACE_LAYER")) : -1; // 1..15: stamps inside that conv
            if (chain_trace && chain_trace_layer >= 1 && chain_trace_layer < NCONV) {
                CU(cudaMemsetAsync(d_ct + 32, 0, 16 * sizeof(unsigned long long), s));
                cp.p[chain_trace_layer - 1].trace = reinterpret_cast<long long *>(d_ct + 32);
            }
            CU(launch_conv_chain(cp, s));
            c->launches++;
            first = NCONV;
            if (chain_trace) {
                unsigned long long h[31];
                CU(cudaStreamSynchronize(s));
                CU(cudaMemcpy(h, d_ct, sizeof h, cudaMemcpyDeviceToHost));
                fprintf(stderr, "chain trace (ns; layer / barrier):");
                for (int k = 1; k < 31; k++) fprintf(stderr, " %llu%s", h[k] - h[k - 1], (k & 1) ? "/" : "");
                fprintf(stderr, "  total %llu\n", h[30] - h[0]);
                if (chain_trace_layer >= 1) {
                    unsigned long long g[8];
                    CU(cudaMemcpy(g, d_ct + 32, sizeof g, cudaMemcpyDeviceToHost));
                    fprintf(stderr, "conv %d stamps (ns since layer entry): init %llu | first A box %llu | first B slot %llu | last commit %llu | acc ready %llu | stored %llu\n",
                            chain_trace_layer, g[1] - g[0], g[2] - g[0], g[3] - g[0], g[4] - g[0], g[5] - g[0], g[6] - g[0]);
                }
            }
        }
        for (int li = first; li < NCONV; li++) {
            ConvParams &p = S.conv_p[li];
            p.nimg = n;
            static const int trace_layer = getenv("MLT_TRACE_LAYER") ? atoi(getenv("MLT_TRACE_LAYER")) : -1;
            static long long *d_trace = nullptr;
            if (li == trace_layer) { // debug: per-tile clock64 stamps of CTA 0's MMA warp -> stderr (last call wins)
                if (!d_trace) CU(cudaMalloc(&d_trace, 1024 * sizeof(long long)));
                CU(cudaMemsetAsync(d_trace, 0, 1024 * sizeof(long long), s));
                p.trace = d_trace;
            }
            // layer3 on a small batch (fewer tiles than half the SMs): 256 output channels split over 4 CTAs per tile
            static const bool no_split = getenv("MLT_NO_SPLIT") != nullptr; // A/B switch for measurements
            const bool split = li >= CONV_SPLIT_FIRST && !no_split && (n + 1) / 2 * 2 <= c->num_sms && li != trace_layer;
            // tiny batches: layer3 8 ways / layer2 4 ways while the items still fit one wave
            const bool split8 = split && (n + 1) / 2 * CONV_SPLIT8_WAYS <= c->num_sms;
            const bool split_l2 = li >= CONV_L2SPLIT_FIRST && li < CONV_SPLIT_FIRST && !no_split && 2 * n * CONV_L2SPLIT_WAYS <= c->num_sms / 2 && li != trace_layer;
            if (split8) {
                ConvParams &q = S.conv_split8[li];
                q.nimg = n;
                CU(launch_conv_umma(li + CONV_SPLIT8_OFFSET, q, c->num_sms, s));
            } else if (split_l2) {
                ConvParams &q = S.conv_split[li];
                q.nimg = n;
                CU(launch_conv_umma(li + CONV_L2SPLIT_OFFSET, q, c->num_sms, s));
            } else if (split) {
                ConvParams &q = S.conv_split[li];
                q.nimg = n;
                CU(launch_conv_umma(li + CONV_SPLIT_OFFSET, q, c->num_sms, s));
            } else
                CU(launch_conv_umma(li, p, c->num_sms, s));
            if (li == trace_layer && getenv("MLT_TRACE_DUMP")) {
                static long long h[1024];
                CU(cudaStreamSynchronize(s));
                CU(cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost));
                fprintf(stderr, "trace layer %d:", li);
                for (int k = 1; k < 1024 && h[k]; k++) fprintf(stderr, " %lld", h[k] - h[k - 1]);
                fprintf(stderr, "\n");
            }
            c->launches++;
            if (prof) CU(cudaEventRecord(c->prof_ev[ev++], s));
        }
        for (int i = 0; i < 3; i++) hp.gap_part[i] = S.gap_part[i];
        CU(launch_head_h(hp, s));
        c->launches++;
        if (prof) { CU(cudaEventRecord(c->prof_ev[ev++], s)); c->prof_valid = true; }
    } else {
        // ---- fp32 CUDA-core cross-check engine (tests)
        int rc = ensure_f32_buffers(c);
        if (rc) return rc;
        CU(launch_stage_conv1_f(ctus, n, secp<float>(c, SEC_CONV1_F32), c->act_f[0], s));
        c->launches++;
        for (int li = 0; li < NCONV; li++) {
            const LayerDesc &L = kLayers[li];
            const bool conv2 = (li & 1) != 0;
            const float *x_block = c->act_f[li & ~1];
            const float *res = nullptr;
            if (conv2) {
                if (L.sc >= 0) {
                    const LayerDesc &L1 = kLayers[li - 1];
                    CU(launch_conv_simt(x_block, secp<float>(c, SEC_SC_W_F32 + L.sc), secp<float>(c, SEC_SC_BIAS + L.sc), nullptr,
                                        c->scratch_f, n, L1.hout * L1.stride, L1.cin, L.cout, 1, 2, 0, s));
                    c->launches++;
                    res = c->scratch_f;
                } else res = x_block;
            }
            CU(launch_conv_simt(c->act_f[li], secp<float>(c, SEC_W_F32 + li), secp<float>(c, SEC_BIAS + li), res, c->act_f[li + 1], n,
                                L.hout * L.stride, L.cin, L.cout, 3, L.stride, 1, s));
            c->launches++;
        }
        hp.act[0] = c->act_f[8]; hp.act[1] = c->act_f[12]; hp.act[2] = c->act_f[16];
        CU(launch_head_f(hp, s));
        c->launches++;
    }
    c->last_n = n;
    return MLT_OK;
}

// gather one CTU (two strided 128x128 int16 blocks) into the dense pinned staging buffer
void gather_ctu(int16_t *dst, const int16_t *org, int org_stride, const int16_t *pred, int pred_stride)
{
    for (int y = 0; y < CTU; y++) memcpy(dst + (size_t)y * CTU, org + (size_t)y * org_stride, CTU * sizeof(int16_t));
    dst += (size_t)CTU * CTU;
    for (int y = 0; y < CTU; y++) memcpy(dst + (size_t)y * CTU, pred + (size_t)y * pred_stride, CTU * sizeof(int16_t));
}

void dense_descs(mlt_ctx *c, int n, const int32_t *pocqp, const mlt_ctu_desc *descs, int slot = 0)
{
    const mlt_ctx::HostSlot &S = c->slot[slot];
    for (int i = 0; i < n; i++) {
        CtuDev &d = S.h_ctus[i];
        d.org = S.d_in + (size_t)i * CTU_IN_ELEMS;
        d.pred = d.org + (size_t)CTU * CTU;
        d.org_stride = d.pred_stride = CTU;
        d.poc = descs ? descs[i].poc : pocqp[2 * i];
        d.qp = descs ? descs[i].qp : pocqp[2 * i + 1];
    }
}

// h_in[0..n) and h_ctus[0..n) are filled: upload, run, download, wait.
int run_host_batch(mlt_ctx *c, int n, mlt_result *out, bool upload_in)
{
    cudaStream_t s = c->stream;
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "submitted batches must be collected first");
    if (upload_in) CU(cudaMemcpyAsync(c->d_in, c->h_in, (size_t)n * CTU_IN_ELEMS * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->d_ctus, c->h_ctus, (size_t)n * sizeof(CtuDev), cudaMemcpyHostToDevice, s));
    int rc = run_network(c, c->d_ctus, n, c->d_out, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->h_out, c->d_out, (size_t)n * sizeof(mlt_result), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(out, c->h_out, (size_t)n * sizeof(mlt_result));
    return MLT_OK;
}

// Host batch of n CTUs, pipelined: chunk i+1 is staged / copied (copy stream) while chunk i computes (compute stream).
// `src` != nullptr: dense caller buffer, copied straight from it (pinned or pageable, no intermediate host copy);
// otherwise `descs` are gathered chunk by chunk into the pinned staging buffer first.
// Enqueues everything (descriptor upload, chunked H2D, kernels, D2H of the results into the slot's pinned buffer, `done`
// event) without waiting for it.
int enqueue_host_batch(mlt_ctx *c, int slot, int n, const int16_t *src, const mlt_ctu_desc *descs, const int32_t *pocqp,
                       const uint8_t *packed = nullptr)
{
    cudaStream_t s = c->stream;
    mlt_ctx::HostSlot &H = c->slot[slot];
    dense_descs(c, n, pocqp, descs, slot);
    CU(cudaMemcpyAsync(H.d_ctus, H.h_ctus, (size_t)n * sizeof(CtuDev), cudaMemcpyHostToDevice, s));
    // Chunk schedule: only the FIRST chunk's H2D is exposed (every later copy hides behind the previous chunks' kernels,
    // which are slower than PCIe), so the chunks start small (n / 16) and grow by 1.5x up to n / 4.
    int sizes[mlt_ctx::MAX_CHUNKS], nchunks = 0;
    // A pipelined batch submitted while another one is still in flight hides its whole copy under that batch's kernels:
    // two equal chunks (the two activation sets / compute streams, like a device-resident batch), no growing schedule.
    static const bool no_hidden = getenv("MLT_NO_HIDDEN") != nullptr; // A/B switch for measurements
    const bool hidden = c->submitted != c->collected && !no_hidden;
    if (n < 1024 && !(hidden && n >= 2 * 240)) sizes[nchunks++] = n;
    else if (hidden && n / 2 <= c->set[1].cap) {
        sizes[nchunks++] = n - n / 2;
        sizes[nchunks++] = n / 2;
    } else {
        const int cap = n / 4 < c->set[1].cap ? n / 4 : c->set[1].cap;
        for (int rem = n, cur = n / 16 > 120 ? n / 16 : 120; rem > 0;) {
            int m = rem < cur ? rem : cur;
            if (rem - m < cur / 2 && rem <= cap) m = rem; // no tiny tail chunk
            if (nchunks == mlt_ctx::MAX_CHUNKS - 1) m = rem;
            sizes[nchunks++] = m;
            rem -= m;
            cur = cur * 3 / 2 < cap ? cur * 3 / 2 : cap;
        }
    }
    if (const char *ov = getenv("MLT_CHUNKS")) { // measurement override: comma-separated chunk sizes summing to n
        int tmp[mlt_ctx::MAX_CHUNKS], k = 0, sum = 0;
        for (const char *q = ov; *q && k < mlt_ctx::MAX_CHUNKS;) {
            tmp[k] = atoi(q);
            sum += tmp[k++];
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        bool ok = sum == n;
        for (int i = 0; i < k; i++) ok = ok && tmp[i] > 0 && tmp[i] <= c->set[i & 1].cap;
        if (ok) { nchunks = k; for (int i = 0; i < k; i++) sizes[i] = tmp[i]; }
    }
    int per = 0;
    for (int i = 0; i < nchunks; i++) per = sizes[i] > per ? sizes[i] : per;
    // chunks alternate between the two activation sets / compute streams (product engines only: the fp32 cross-check
    // engine has ONE set of fp32 buffers, its chunks must stay on one stream)
    const bool two = nchunks > 1 && !c->profiling && c->engine != 1;
    if (two) { CU(cudaEventRecord(c->ev_fork, s)); CU(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0)); } // descriptors uploaded
    for (int i = 0, off = 0; off < n; off += sizes[i], i++) {
        const int m = sizes[i];
        const int16_t *from = src ? src + (size_t)off * CTU_IN_ELEMS : c->h_in + (size_t)off * CTU_IN_ELEMS;
        if (packed) {
            CU(cudaMemcpyAsync(H.d_packed + (size_t)off * CTU_PACKED_BYTES, packed + (size_t)off * CTU_PACKED_BYTES, (size_t)m * CTU_PACKED_BYTES,
                               cudaMemcpyHostToDevice, c->copy_stream));
        } else {
            if (!src)
                for (int k = off; k < off + m; k++)
                    gather_ctu(c->h_in + (size_t)k * CTU_IN_ELEMS, descs[k].org, descs[k].org_stride, descs[k].pred, descs[k].pred_stride);
            CU(cudaMemcpyAsync(H.d_in + (size_t)off * CTU_IN_ELEMS, from, (size_t)m * CTU_IN_ELEMS * sizeof(int16_t),
                               cudaMemcpyHostToDevice, c->copy_stream));
        }
        CU(cudaEventRecord(c->ev_in[i], c->copy_stream));
        const int si = two ? (i & 1) : 0;
        cudaStream_t cs = si ? c->stream2 : s;
        CU(cudaStreamWaitEvent(cs, c->ev_in[i], 0));
        if (packed) { // 40 KiB -> 64 KiB per CTU on the chunk's compute stream, right in front of its stem kernel
            CU(launch_unpack10(H.d_packed + (size_t)off * CTU_PACKED_BYTES, H.d_in + (size_t)off * CTU_IN_ELEMS, (size_t)m * CTU_IN_ELEMS, cs));
            c->launches++;
        }
        const int rc = run_network(c, H.d_ctus + off, m, H.d_out + off, cs, si);
        if (rc) { cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->stream2); return rc; }
    }
    if (two) { CU(cudaEventRecord(c->ev_join, c->stream2)); CU(cudaStreamWaitEvent(s, c->ev_join, 0)); }
    c->last_n = n <= per ? n : 0; // debug_activation only sees a whole batch when it ran as one chunk
    CU(cudaMemcpyAsync(H.h_out, H.d_out, (size_t)n * sizeof(mlt_result), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(H.done, s));
    H.n = n;
    return MLT_OK;
}

int ensure_packed(mlt_ctx *c, mlt_ctx::HostSlot &H)
{
    if (!H.d_packed) CU(cudaMalloc(&H.d_packed, (size_t)c->max_batch * CTU_PACKED_BYTES));
    return MLT_OK;
}

int run_host_batch_chunked(mlt_ctx *c, int n, const int16_t *src, const mlt_ctu_desc *descs, const int32_t *pocqp, mlt_result *out,
                           const uint8_t *packed = nullptr)
{
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "%llu submitted batch(es) not collected yet", (unsigned long long)(c->submitted - c->collected));
    const int rc = enqueue_host_batch(c, 0, n, src, descs, pocqp, packed);
    if (rc) return rc;
    CU(cudaEventSynchronize(c->slot[0].done));
    memcpy(out, c->slot[0].h_out, (size_t)n * sizeof(mlt_result));
    return MLT_OK;
}

int check_ctx(mlt_ctx *c, bool device_entry = false)
{
    if (!c) return MLT_E_INVAL;
    c->err.clear();
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, MLT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (!device_entry && c->dev_pending_host) {
        // a device-resident call may still be running on the caller's stream: every stream of the context waits for it
        for (cudaStream_t st : {c->stream, c->stream2, c->copy_stream}) CU(cudaStreamWaitEvent(st, c->ev_dev, 0));
        c->dev_pending_host = false;
    }
    return MLT_OK;
}

} // namespace

// ================================================================================================ C ABI

extern "C" {

int mlt_abi_version(void) { return MLT_ABI_VERSION; }

const char *mlt_strerror(int rc)
{
    switch (rc) {
    case MLT_OK: return "ok";
    case MLT_E_INVAL: return "invalid argument";
    case MLT_E_IO: return "weight file unreadable";
    case MLT_E_FORMAT: return "weight file is not an MLTW blob for this architecture";
    case MLT_E_CUDA: return "CUDA error";
    case MLT_E_NOMEM: return "out of memory";
    case MLT_E_NODEVICE: return "no sm_100 CUDA device";
    case MLT_E_BATCH: return "batch larger than max_batch";
    case MLT_E_STATE: return "call sequence error";
    default: return "unknown error";
    }
}

const char *mlt_last_error(const mlt_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

void mlt_destroy(mlt_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &S : c->set) { for (int a = 0; a < NACT; a++) cudaFree(S.act_h[a]); cudaFree(S.act0q); for (float *g : S.gap_part) cudaFree(g); }
    for (int a = 0; a < NACT; a++) cudaFree(c->act_f[a]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_dev) cudaEventDestroy(c->ev_dev);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    for (auto &H : c->slot) cudaFree(H.d_packed);
    if (c->slot[1].d_in) { cudaFree(c->slot[1].d_in); cudaFree(c->slot[1].d_ctus); cudaFree(c->slot[1].d_out); cudaFreeHost(c->slot[1].h_ctus); cudaFreeHost(c->slot[1].h_out); }
    for (auto &H : c->slot) if (H.done) cudaEventDestroy(H.done);
    cudaFree(c->scratch_f); cudaFree(c->d_blob); cudaFree(c->d_in); cudaFree(c->d_ctus); cudaFree(c->d_out);
    cudaFree(c->d_dbg); cudaFree(c->d_pic); cudaFree(c->d_ref); cudaFree(c->d_pic_ctus);
    cudaFree(c->d_me_cost); cudaFree(c->d_me_best); cudaFree(c->d_me_mv); cudaFreeHost(c->h_me_best); cudaFreeHost(c->h_me_mv);
    cudaFreeHost(c->h_pic_ctus); cudaFreeHost(c->h_in); cudaFreeHost(c->h_ctus); cudaFreeHost(c->h_out);
    for (cudaEvent_t e : c->prof_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_in) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int mlt_create_ex(mlt_ctx **out, const char *weights_path, int cuda_device, int max_batch)
{
    if (!out) return MLT_E_INVAL;
    *out = nullptr;
    if (!weights_path || max_batch < 1 || max_batch > (1 << 16)) return MLT_E_INVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MLT_E_NODEVICE;
    if (cuda_device < 0 || cuda_device >= ndev) return MLT_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cuda_device) != cudaSuccess) return MLT_E_NODEVICE;
    if (prop.major != 10) return MLT_E_NODEVICE; // sm_100a cubin only: tcgen05 / TMEM required, no other path exists
    mlt_ctx *c = new (std::nothrow) mlt_ctx();
    if (!c) return MLT_E_NOMEM;
    c->device = cuda_device;
    c->max_batch = max_batch;
    c->num_sms = prop.multiProcessorCount;
    int rc = MLT_OK;
    auto body = [&]() -> int {
        CU(cudaSetDevice(cuda_device));
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : c->ev_in) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        int r = load_blob(c, weights_path);
        if (r) return r;
        CU(conv_umma_init());
        CU(stem_umma_init());
        CU(stem5_umma_init());
        CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_dev, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        const ActLayout l0q{64, 32, 0, 0};
        for (int si = 0; si < 2; si++) {
            mlt_ctx::ActSet &S = c->set[si];
            S.cap = si == 0 ? max_batch : (max_batch + 1) / 2;
            for (int a = 1; a < NACT; a++) {
                // zeroed once: the unused half of the last image pair of an odd batch must stay finite
                const ActLayout L = act_layout(a);
                const size_t bytes = L.unit_elems() * L.units_for(S.cap) * sizeof(__half);
                CU(cudaMalloc(&S.act_h[a], bytes));
                CU(cudaMemsetAsync(S.act_h[a], 0, bytes, c->stream));
            }
            CU(cudaMalloc(&S.act0q, l0q.unit_elems() * (size_t)S.cap * sizeof(__half)));
            static const int gp_per_img[3] = {8 * 4 * 64, 2 * 4 * 128, 4 * 256};
            for (int i = 0; i < 3; i++) CU(cudaMalloc(&S.gap_part[i], ((size_t)S.cap + 1) * gp_per_img[i] * sizeof(float)));
            memset(S.conv_p, 0, sizeof S.conv_p);
            for (int li = 1; li < NCONV; li++) { // conv 0 runs inside the stem kernel (ensure_act0 prepares it for the unfused engine)
                ConvParams &p = S.conv_p[li];
                const bool conv2 = (li & 1) != 0;
                const int xa = li & ~1; // input of the BasicBlock this conv belongs to
                const ActLayout in_l = act_layout(li), x_l = xa == 0 ? l0q : act_layout(xa);
                const __half *xp = xa == 0 ? S.act0q : S.act_h[xa];
                CU(conv_umma_prepare(li, &p, S.act_h[li], in_l, conv2 ? xp : nullptr, conv2 ? &x_l : nullptr, (size_t)S.cap));
                p.w = secp<__half>(c, SEC_W_F16 + li);
                p.bias = secp<__half>(c, SEC_BIAS_MMA + li);
                p.bias_f32 = secp<float>(c, SEC_BIAS_FUSED + li);
                p.x_w = conv2 ? secp<__half>(c, SEC_X_W_F16 + li) : nullptr;
                p.out = S.act_h[li + 1];
                p.gap_part = li == 7 ? S.gap_part[0] : (li == 11 ? S.gap_part[1] : (li == 15 ? S.gap_part[2] : nullptr));
                p.relu = 1;
                p.dbg = getenv("MLT_DEBUG_FLAGS") ? atoi(getenv("MLT_DEBUG_FLAGS")) : 0;
                p.reverse = getenv("MLT_NO_REVERSE") ? 0 : (li & 1); // the stem walks the images upwards, conv 1 downwards, conv 2 upwards, ...
                if (li >= CONV_L2SPLIT_FIRST) {
                    S.conv_split[li] = p;
                    S.conv_split[li].w = secp<__half>(c, SEC_W_SPLIT + li);
                    S.conv_split[li].x_w = conv2 ? secp<__half>(c, SEC_X_SPLIT + li) : nullptr;
                }
                if (li >= CONV_SPLIT_FIRST) {
                    S.conv_split8[li] = p;
                    S.conv_split8[li].w = secp<__half>(c, SEC_W_SPLIT8 + li);
                    S.conv_split8[li].x_w = conv2 ? secp<__half>(c, SEC_X_SPLIT8 + li) : nullptr;
                }
            }
        }
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMalloc(&c->d_in, (size_t)max_batch * CTU_IN_ELEMS * sizeof(int16_t)));
        CU(cudaMalloc(&c->d_ctus, (size_t)max_batch * sizeof(CtuDev)));
        CU(cudaMalloc(&c->d_out, (size_t)max_batch * sizeof(mlt_result)));
        CU(cudaHostAlloc(&c->h_in, (size_t)max_batch * CTU_IN_ELEMS * sizeof(int16_t), cudaHostAllocDefault));
        CU(cudaHostAlloc(&c->h_ctus, (size_t)max_batch * sizeof(CtuDev), cudaHostAllocDefault));
        CU(cudaHostAlloc(&c->h_out, (size_t)max_batch * sizeof(mlt_result), cudaHostAllocDefault));
        c->slot[0].d_in = c->d_in; c->slot[0].d_ctus = c->d_ctus; c->slot[0].h_ctus = c->h_ctus;
        c->slot[0].d_out = c->d_out; c->slot[0].h_out = c->h_out;
        for (auto &H : c->slot) CU(cudaEventCreateWithFlags(&H.done, cudaEventDisableTiming));
        return MLT_OK;
    };
    rc = body();
    if (rc != MLT_OK) {
        fprintf(stderr, "mlt_create: %s (%s)\n", mlt_strerror(rc), c->err.c_str());
        mlt_destroy(c);
        return rc;
    }
    *out = c;
    return MLT_OK;
}

int mlt_create(mlt_ctx **out, const char *weights_path, int cuda_device)
{
    return mlt_create_ex(out, weights_path, cuda_device, 512); // 480 eligible CTUs in a 2160p picture
}

int mlt_set_engine(mlt_ctx *c, int engine)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (engine < 0 || engine > 2) return fail(c, MLT_E_INVAL, "engine must be 0 (tcgen05), 1 (fp32 cross-check) or 2 (tcgen05, unfused stem)");
    c->engine = engine;
    return MLT_OK;
}

uint64_t mlt_launch_count(const mlt_ctx *c) { return c ? c->launches : 0; }

int mlt_set_profiling(mlt_ctx *c, int on)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (on && !c->prof_ev[0])
        for (cudaEvent_t &e : c->prof_ev) CU(cudaEventCreate(&e));
    c->profiling = on != 0;
    c->prof_valid = false;
    return MLT_OK;
}

int mlt_get_profile(mlt_ctx *c, float *ms, int capacity)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    constexpr int NK = NCONV + 2;
    if (!ms || capacity < NK) return fail(c, MLT_E_INVAL, "capacity must be >= %d", NK);
    if (!c->prof_valid) return fail(c, MLT_E_STATE, "no profiled batch has been run");
    CU(cudaEventSynchronize(c->prof_ev[NK]));
    for (int i = 0; i < NK; i++) CU(cudaEventElapsedTime(&ms[i], c->prof_ev[i], c->prof_ev[i + 1]));
    return NK;
}

int mlt_predict_batch(mlt_ctx *c, int n, const mlt_ctu_desc *descs, mlt_result *out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!descs || !out))) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->max_batch);
    if (n == 0) return MLT_OK;
    for (int i = 0; i < n; i++)
        if (!descs[i].org || !descs[i].pred) return fail(c, MLT_E_INVAL, "descs[%d]: null org/pred", i);
    return run_host_batch_chunked(c, n, nullptr, descs, nullptr, out);
}

int mlt_predict_ctu(mlt_ctx *c, const int16_t *org, int org_stride, const int16_t *pred, int pred_stride, int poc, int qp,
                    mlt_result *out)
{
    mlt_ctu_desc d;
    d.org = org; d.pred = pred; d.org_stride = org_stride; d.pred_stride = pred_stride; d.poc = poc; d.qp = qp;
    return mlt_predict_batch(c, 1, &d, out);
}

int mlt_predict_batch_dense(mlt_ctx *c, int n, const int16_t *orgpred, const int32_t *pocqp, mlt_result *out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!orgpred || !pocqp || !out))) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->max_batch);
    if (n == 0) return MLT_OK;
    return run_host_batch_chunked(c, n, orgpred, nullptr, pocqp, out);
}

static int submit_host_batch(mlt_ctx *c, int n, const int16_t *orgpred, const uint8_t *packed, const int32_t *pocqp)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n < 1 || (!orgpred && !packed) || !pocqp) return fail(c, MLT_E_INVAL, "null argument / empty batch");
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->max_batch);
    if (c->submitted - c->collected >= 2) return fail(c, MLT_E_STATE, "two batches already in flight: call mlt_collect first");
    const int si = (int)(c->submitted & 1);
    mlt_ctx::HostSlot &H = c->slot[si];
    if (!H.d_in) { // second slot: allocated on first use
        CU(cudaMalloc(&H.d_in, (size_t)c->max_batch * CTU_IN_ELEMS * sizeof(int16_t)));
        CU(cudaMalloc(&H.d_ctus, (size_t)c->max_batch * sizeof(CtuDev)));
        CU(cudaMalloc(&H.d_out, (size_t)c->max_batch * sizeof(mlt_result)));
        CU(cudaHostAlloc(&H.h_ctus, (size_t)c->max_batch * sizeof(CtuDev), cudaHostAllocDefault));
        CU(cudaHostAlloc(&H.h_out, (size_t)c->max_batch * sizeof(mlt_result), cudaHostAllocDefault));
    }
    if (packed && (rc = ensure_packed(c, H)) != MLT_OK) return rc;
    rc = enqueue_host_batch(c, si, n, orgpred, nullptr, pocqp, packed);
    if (rc) return rc;
    H.busy = true;
    c->submitted++;
    return MLT_OK;
}

int mlt_submit_batch_dense(mlt_ctx *c, int n, const int16_t *orgpred, const int32_t *pocqp) { return submit_host_batch(c, n, orgpred, nullptr, pocqp); }

int mlt_submit_batch_packed10(mlt_ctx *c, int n, const uint8_t *packed, const int32_t *pocqp) { return submit_host_batch(c, n, nullptr, packed, pocqp); }

int mlt_predict_batch_packed10(mlt_ctx *c, int n, const uint8_t *packed, const int32_t *pocqp, mlt_result *out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!packed || !pocqp || !out))) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->max_batch);
    if (n == 0) return MLT_OK;
    if ((rc = ensure_packed(c, c->slot[0])) != MLT_OK) return rc;
    return run_host_batch_chunked(c, n, nullptr, nullptr, pocqp, out, packed);
}

int mlt_collect(mlt_ctx *c, mlt_result *out, int *n_out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!out) return fail(c, MLT_E_INVAL, "null argument");
    if (c->submitted == c->collected) return fail(c, MLT_E_STATE, "no submitted batch to collect");
    mlt_ctx::HostSlot &H = c->slot[c->collected & 1];
    CU(cudaEventSynchronize(H.done));
    memcpy(out, H.h_out, (size_t)H.n * sizeof(mlt_result));
    if (n_out) *n_out = H.n;
    H.busy = false;
    c->collected++;
    return MLT_OK;
}

int mlt_predict_batch_device(mlt_ctx *c, int n, const int16_t *d_orgpred, const int32_t *d_pocqp, mlt_result *d_out,
                             void *cuda_stream)
{
    int rc = check_ctx(c, true);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_orgpred || !d_pocqp || !d_out))) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->max_batch);
    if (((uintptr_t)d_orgpred & 15) != 0) return fail(c, MLT_E_INVAL, "d_orgpred must be 16-byte aligned");
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "submitted batches must be collected first");
    if (n == 0) return MLT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    // order this call after the previous device call (possibly on another stream): both use the context's buffers
    if (c->dev_issued) CU(cudaStreamWaitEvent(s, c->ev_dev, 0));
    struct Mark { // ... and leave the marker for whoever comes next, on every exit path
        mlt_ctx *c; cudaStream_t s;
        ~Mark() { if (cudaEventRecord(c->ev_dev, s) == cudaSuccess) c->dev_issued = c->dev_pending_host = true; }
    } mark{c, s};
    // descriptors are built on the device: nothing crosses PCIe on this path
    dense_descs_kernel<<<(n + 255) / 256, 256, 0, s>>>(c->d_ctus, d_orgpred, d_pocqp, n);
    CU(cudaGetLastError());
    c->launches++;
    static const bool no_slice = getenv("MLT_NO_SLICE") != nullptr; // A/B switch for measurements
    if (n < 2 * 240 || c->profiling || c->engine != 0 || no_slice) return run_network(c, c->d_ctus, n, d_out, s);
    // two slices on two streams: the inter-kernel gaps of one slice are filled by the other slice's kernels
    static const int nslices = getenv("MLT_SLICES") ? atoi(getenv("MLT_SLICES")) : 2; // measurement override (even, >= 2)
    CU(cudaEventRecord(c->ev_fork, s));
    CU(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    const int per = (n + nslices - 1) / nslices;
    for (int i = 0, off = 0; off < n; i++, off += per) {
        const int m = n - off < per ? n - off : per;
        rc = run_network(c, c->d_ctus + off, m, d_out + off, (i & 1) ? c->stream2 : s, i & 1);
        if (rc) return rc;
    }
    CU(cudaEventRecord(c->ev_join, c->stream2));
    CU(cudaStreamWaitEvent(s, c->ev_join, 0));
    c->last_n = 0; // debug_activation only sees single-slice batches
    return MLT_OK;
}

int mlt_begin_picture(mlt_ctx *c, const int16_t *org_luma, int stride, int width, int height, int poc)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!org_luma || width < CTU || height < CTU || stride < width) return fail(c, MLT_E_INVAL, "bad picture geometry");
    const int pitch = (width + 7) & ~7; // rows 16-byte aligned on the device
    const size_t need = (size_t)pitch * height;
    if (need > c->pic_capacity) {
        if (c->d_pic) cudaFree(c->d_pic);
        c->d_pic = nullptr;
        c->pic_capacity = 0;
        c->pic_valid = false;
        CU(cudaMalloc(&c->d_pic, need * sizeof(int16_t)));
        c->pic_capacity = need;
    }
    CU(cudaMemcpy2DAsync(c->d_pic, (size_t)pitch * sizeof(int16_t), org_luma, (size_t)stride * sizeof(int16_t),
                         (size_t)width * sizeof(int16_t), height, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream)); // the caller may reuse / modify its buffer after return
    c->pic_pitch = pitch; c->pic_w = width; c->pic_h = height; c->pic_poc = poc; c->pic_valid = true;
    c->ref_valid = false; // a reference plane belongs to the picture it was uploaded for
    return MLT_OK;
}

int mlt_predict_ctu_in_picture(mlt_ctx *c, int x, int y, const int16_t *pred, int pred_stride, int qp, mlt_result *out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!c->pic_valid) return fail(c, MLT_E_STATE, "mlt_begin_picture has not been called");
    if (!pred || !out) return fail(c, MLT_E_INVAL, "null argument");
    // same gate as EncCu.cpp:755: the CTU must lie fully inside the picture; x must keep 16-byte alignment
    if (x < 0 || y < 0 || x + CTU > c->pic_w || y + CTU > c->pic_h || (x & 7)) return fail(c, MLT_E_INVAL, "CTU (%d,%d) not inside the picture", x, y);
    int16_t *hp = c->h_in + (size_t)CTU * CTU; // pred plane slot of CTU 0
    for (int r = 0; r < CTU; r++) memcpy(hp + (size_t)r * CTU, pred + (size_t)r * pred_stride, CTU * sizeof(int16_t));
    cudaStream_t s = c->stream;
    CU(cudaMemcpyAsync(c->d_in + (size_t)CTU * CTU, hp, (size_t)CTU * CTU * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    CtuDev &d = c->h_ctus[0];
    d.org = c->d_pic + (size_t)y * c->pic_pitch + x;
    d.org_stride = c->pic_pitch;
    d.pred = c->d_in + (size_t)CTU * CTU;
    d.pred_stride = CTU;
    d.poc = c->pic_poc;
    d.qp = qp;
    return run_host_batch(c, 1, out, false);
}

int mlt_pin_host_buffer(mlt_ctx *c, const void *ptr, uint64_t bytes)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!ptr || !bytes) return fail(c, MLT_E_INVAL, "null buffer");
    const cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return MLT_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, MLT_E_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    return MLT_OK;
}

int mlt_unpin_host_buffer(mlt_ctx *c, const void *ptr)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!ptr) return fail(c, MLT_E_INVAL, "null buffer");
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "submitted batches may still be reading host buffers: collect them first");
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    const cudaError_t e = cudaHostUnregister(const_cast<void *>(ptr));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, MLT_E_INVAL, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return MLT_OK;
}

int mlt_picture_ctu_count(const mlt_ctx *c)
{
    if (!c) return MLT_E_INVAL;
    if (!c->pic_valid) return MLT_E_STATE;
    return (c->pic_w / CTU) * (c->pic_h / CTU); // CTUs lying fully inside the picture (EncCu.cpp:755)
}

// upload the reference plane of the picture begun (same pitch as the org plane) and make sure the CTU list buffers exist
static int upload_reference(mlt_ctx *c, const int16_t *ref_luma, int ref_stride)
{
    const size_t need = (size_t)c->pic_pitch * c->pic_h;
    if (need > c->ref_capacity) {
        if (c->d_ref) cudaFree(c->d_ref);
        c->d_ref = nullptr;
        c->ref_capacity = 0;
        CU(cudaMalloc(&c->d_ref, need * sizeof(int16_t)));
        c->ref_capacity = need;
    }
    c->ref_valid = false;
    CU(cudaMemcpy2DAsync(c->d_ref, (size_t)c->pic_pitch * sizeof(int16_t), ref_luma, (size_t)ref_stride * sizeof(int16_t),
                         (size_t)c->pic_w * sizeof(int16_t), c->pic_h, cudaMemcpyHostToDevice, c->stream));
    c->ref_valid = true;
    return MLT_OK;
}

static int ensure_pic_ctus(mlt_ctx *c)
{
    if (!c->d_pic_ctus) {
        CU(cudaMalloc(&c->d_pic_ctus, (size_t)c->max_batch * sizeof(PicCtu)));
        CU(cudaMallocHost(&c->h_pic_ctus, (size_t)c->max_batch * sizeof(PicCtu)));
    }
    return MLT_OK;
}

int mlt_estimate_picture_mv(mlt_ctx *c, const int16_t *ref_luma, int ref_stride, int range, int16_t *mv_out, uint32_t *cost_out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!c->pic_valid) return fail(c, MLT_E_STATE, "mlt_begin_picture has not been called");
    if (!mv_out || range < 0 || range > MLT_ME_MAX_RANGE) return fail(c, MLT_E_INVAL, "bad argument (range 0..%d)", MLT_ME_MAX_RANGE);
    if (ref_luma ? ref_stride < c->pic_w : !c->ref_valid) return fail(c, ref_luma ? MLT_E_INVAL : MLT_E_STATE, "no reference picture");
    const int cols = c->pic_w / CTU, n = cols * (c->pic_h / CTU);
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "picture has %d eligible CTUs > max_batch=%d", n, c->max_batch);
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "submitted batches must be collected first");
    cudaStream_t s = c->stream;
    if (ref_luma && (rc = upload_reference(c, ref_luma, ref_stride)) != MLT_OK) return rc;
    if ((rc = ensure_pic_ctus(c)) != MLT_OK) return rc;
    if (!c->d_me_cost) {
        constexpr int DMAX = 2 * MLT_ME_MAX_RANGE + 1;
        CU(cudaMalloc(&c->d_me_cost, (size_t)c->max_batch * DMAX * DMAX * sizeof(unsigned)));
        CU(cudaMalloc(&c->d_me_best, (size_t)c->max_batch * sizeof(unsigned)));
        CU(cudaMalloc(&c->d_me_mv, (size_t)c->max_batch * 2 * sizeof(int16_t)));
        CU(cudaMallocHost(&c->h_me_best, (size_t)c->max_batch * sizeof(unsigned)));
        CU(cudaMallocHost(&c->h_me_mv, (size_t)c->max_batch * 2 * sizeof(int16_t)));
    }
    for (int i = 0; i < n; i++) c->h_pic_ctus[i] = PicCtu{(i % cols) * CTU, (i / cols) * CTU, 0, 0};
    CU(cudaMemcpyAsync(c->d_pic_ctus, c->h_pic_ctus, (size_t)n * sizeof(PicCtu), cudaMemcpyHostToDevice, s));
    CU(launch_picture_me(c->d_pic, c->d_ref, c->pic_pitch, c->pic_w, c->pic_h, c->d_pic_ctus, n, range, c->d_me_cost, c->d_me_mv, c->d_me_best, s));
    c->launches += 2;
    CU(cudaMemcpyAsync(c->h_me_mv, c->d_me_mv, (size_t)n * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(c->h_me_best, c->d_me_best, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(mv_out, c->h_me_mv, (size_t)n * 2 * sizeof(int16_t));
    if (cost_out) memcpy(cost_out, c->h_me_best, (size_t)n * sizeof(uint32_t));
    return n;
}

int mlt_predict_picture(mlt_ctx *c, const int16_t *ref_luma, int ref_stride, const int16_t *mv, const int32_t *ctu_qp, int slice_qp,
                        mlt_result *out, int capacity)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!c->pic_valid) return fail(c, MLT_E_STATE, "mlt_begin_picture has not been called");
    if (!out || (ref_luma && ref_stride < c->pic_w)) return fail(c, MLT_E_INVAL, "bad reference picture");
    if (!ref_luma && !c->ref_valid) return fail(c, MLT_E_STATE, "no reference picture uploaded for this picture");
    const int cols = c->pic_w / CTU, n = cols * (c->pic_h / CTU);
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "picture has %d eligible CTUs > max_batch=%d", n, c->max_batch);
    if (capacity < n) return fail(c, MLT_E_INVAL, "out holds %d results, the picture has %d eligible CTUs", capacity, n);
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "submitted batches must be collected first");
    cudaStream_t s = c->stream;
    if (ref_luma && (rc = upload_reference(c, ref_luma, ref_stride)) != MLT_OK) return rc;
    if ((rc = ensure_pic_ctus(c)) != MLT_OK) return rc;
    for (int i = 0; i < n; i++) {
        const int x = (i % cols) * CTU, y = (i / cols) * CTU;
        c->h_pic_ctus[i] = PicCtu{x, y, mv ? mv[2 * i] : 0, mv ? mv[2 * i + 1] : 0};
        CtuDev &d = c->h_ctus[i];
        d.org = c->d_pic + (size_t)y * c->pic_pitch + x; // read in place: CTU columns are 256-byte aligned
        d.org_stride = c->pic_pitch;
        d.pred = c->d_in + (size_t)i * CTU_IN_ELEMS + (size_t)CTU * CTU;
        d.pred_stride = CTU;
        d.poc = c->pic_poc;
        d.qp = ctu_qp ? ctu_qp[i] : slice_qp;
    }
    CU(cudaMemcpyAsync(c->d_pic_ctus, c->h_pic_ctus, (size_t)n * sizeof(PicCtu), cudaMemcpyHostToDevice, s));
    CU(launch_picture_pred(c->d_ref, c->pic_pitch, c->pic_w, c->pic_h, c->d_pic_ctus, n, c->d_in, s));
    c->launches++;
    rc = run_host_batch(c, n, out, false);
    return rc ? rc : n;
}

int mlt_debug_picture_pred(mlt_ctx *c, int16_t *out, int capacity)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!c->pic_valid || !c->d_pic_ctus) return fail(c, MLT_E_STATE, "no picture pre-pass has run");
    const int n = (c->pic_w / CTU) * (c->pic_h / CTU);
    if (!out || capacity < n) return fail(c, MLT_E_INVAL, "out too small");
    CU(cudaMemcpy2D(out, (size_t)CTU * CTU * sizeof(int16_t), c->d_in + (size_t)CTU * CTU, CTU_IN_ELEMS * sizeof(int16_t),
                    (size_t)CTU * CTU * sizeof(int16_t), n, cudaMemcpyDeviceToHost));
    return n;
}

int mlt_debug_stage(mlt_ctx *c, int n, const mlt_ctu_desc *descs, float *out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n < 1 || !descs || !out) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->max_batch) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->max_batch);
    for (int i = 0; i < n; i++)
        gather_ctu(c->h_in + (size_t)i * CTU_IN_ELEMS, descs[i].org, descs[i].org_stride, descs[i].pred, descs[i].pred_stride);
    dense_descs(c, n, nullptr, descs);
    const size_t bytes = (size_t)n * CTU_IN_ELEMS * sizeof(float);
    rc = ensure_dbg(c, bytes);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    CU(cudaMemcpyAsync(c->d_in, c->h_in, (size_t)n * CTU_IN_ELEMS * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->d_ctus, c->h_ctus, (size_t)n * sizeof(CtuDev), cudaMemcpyHostToDevice, s));
    CU(launch_stage(c->d_ctus, n, c->d_dbg, s));
    c->launches++;
    CU(cudaMemcpyAsync(out, c->d_dbg, bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return MLT_OK;
}

int64_t mlt_debug_activation(mlt_ctx *c, int layer, float *out, int64_t capacity)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (layer < 0 || layer >= NACT || !out) return fail(c, MLT_E_INVAL, "bad layer");
    if (c->last_n <= 0) return fail(c, MLT_E_STATE, "no batch has been run");
    const size_t elems = act_elems(layer) * c->last_n;
    if ((int64_t)elems > capacity) return fail(c, MLT_E_INVAL, "capacity %lld < %zu", (long long)capacity, elems);
    cudaStream_t s = c->stream;
    if (c->engine != 1) {
        rc = ensure_dbg(c, elems * sizeof(float));
        if (rc) return rc;
        if (layer == 0 && c->engine == 0) {
            // the fused stem never materialises conv1's output: recompute it for the last batch with the standalone kernel
            rc = ensure_act0(c, c->set[0]);
            if (rc) return rc;
            CU(launch_conv1_umma(c->d_ctus, c->last_n, secp<__half>(c, SEC_CONV1_UMMA), c->set[0].act_h[0], s));
        }
        CU(launch_unpack_act(c->set[0].act_h[layer], c->d_dbg, c->last_n, act_layout(layer), s));
        CU(cudaMemcpyAsync(out, c->d_dbg, elems * sizeof(float), cudaMemcpyDeviceToHost, s));
    } else {
        if (!c->act_f[layer]) return fail(c, MLT_E_STATE, "fp32 engine has not run");
        CU(cudaMemcpyAsync(out, c->act_f[layer], elems * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    return (int64_t)elems;
}

} // extern "C"
