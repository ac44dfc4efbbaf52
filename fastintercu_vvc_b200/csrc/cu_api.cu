// cu_api.cu -- C ABI for the smaller-CU models (include/mltcnn_cu.h): context, weight blob, strip buffers, launch
// sequence (staging + conv1, 20 tcgen05 convs, head).  Drop-in for the cuw != 128 branch of the reference hook
// (EncCu.cpp:754,899,916-919).  No CPU fallback.
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
enum : uint32_t { SEC_CONV1_UMMA = 0x002, SEC_STEM_CONV1 = 0x003, SEC_STEM5_W = 0x004, SEC_STEM5_CORR = 0x005, SEC_W_F16 = 0x100, SEC_BIAS_FUSED = 0x200, SEC_FC_W = 0x800, SEC_FC_B = 0x900, SEC_BIAS_MMA = 0xA00, SEC_X_W_F16 = 0xB00 };
constexpr int FC_IN[CU_NHEAD] = {66, 98, 130, 258}, FC_OUT[CU_NHEAD] = {2, 3, 4, 6};

struct Section { const uint8_t *dev = nullptr; size_t bytes = 0; };

} // namespace

struct mlt_cu_ctx {
    int device = 0, size = 0, cap = 0, num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    static constexpr int MAX_CHUNKS = 12;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}; // "chunk i of the input batch has landed in HBM"
    uint8_t *d_blob = nullptr;
    Section sec[0x1000];
    CuLayerInfo info[CU_NCONV];
    ActLayout lay[CU_NACT];
    __half *act[CU_NACT] = {};
    // 64- / 32-px networks: staging + conv1 + layer0.0.conv1 run as the fused stem kernel (stem_umma.cu); conv1's output
    // (act[0]) is then only materialised for debug reads, and `act0q` holds its even / even quarter for layer0.0's shortcut
    bool fused = false;
    __half *act0q = nullptr;
    ActLayout lay0q;
    float *gap_part[CU_NHEAD] = {}; // fp32 pool partial sums of the head-feeding convs on maps >= 8x8
    ConvParams conv_p[CU_NCONV];
    int16_t *d_in = nullptr, *h_in = nullptr; // dense [cap][2][size][size]
    int32_t *d_pq = nullptr, *h_pq = nullptr; // [cap][2]
    CtuDev *d_cus = nullptr;
    mlt_cu_result *d_out = nullptr, *h_out = nullptr;
    // Host-batch slots: slot 0 = the buffers above (every synchronous call); slot 1 is allocated on the first
    // mlt_cu_submit_batch_dense so that two batches can be in flight (the H2D of batch k + 1 runs while batch k computes).
    // The activation strips are shared: all kernels are ordered on the one compute stream.
    struct HostSlot {
        int16_t *d_in = nullptr;
        int32_t *d_pq = nullptr, *h_pq = nullptr;
        mlt_cu_result *d_out = nullptr, *h_out = nullptr;
        cudaEvent_t done = nullptr;
        int n = 0;
        bool busy = false;
    } slot[2];
    uint64_t submitted = 0, collected = 0;
    // end of the last mlt_cu_predict_batch_device (runs on the CALLER's stream, uses the context's buffers): see api.cu
    cudaEvent_t ev_dev = nullptr;
    bool dev_issued = false, dev_pending_host = false;
    int16_t *d_pic = nullptr, *d_mv = nullptr; // mlt_cu_predict_picture: org + reference luma planes (same pitch), per-CU MVs
    size_t pic_capacity = 0;
    float *d_dbg = nullptr;
    size_t dbg_bytes = 0;
    int last_n = 0;
    uint64_t launches = 0;
    std::string err;
};

namespace {

int fail(mlt_cu_ctx *c, int rc, const char *fmt, ...)
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

template <typename T>
const T *secp(const mlt_cu_ctx *c, uint32_t id) { return reinterpret_cast<const T *>(c->sec[id].dev); }

int load_blob(mlt_cu_ctx *c, const char *path)
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
    // the operand layouts depend on the CU size (tile shapes differ per size), so a blob is packed for ONE size
    if (hdr[0] != MLTW_MAGIC || hdr[1] != 2 || (int)hdr[2] != c->size || total != raw.size() || 32 + (size_t)hdr[3] * 24 > raw.size())
        return fail(c, MLT_E_FORMAT, "'%s' is not an MLTW v2 blob for the %dx%d CU model", path, c->size, c->size);
    CU(cudaMalloc(&c->d_blob, raw.size()));
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
    auto need = [&](uint32_t id, size_t bytes) { return c->sec[id].dev != nullptr && c->sec[id].bytes == bytes; };
    bool ok = need(SEC_CONV1_UMMA, 2 * 4 * 32 * 16) && need(SEC_STEM_CONV1, 4 * 1024) && need(SEC_STEM5_W, 7 * 1024) &&
              need(SEC_STEM5_CORR, (2 * 5 * 2 * 32 + 2 * 32 + 32) * 4);
    for (int li = 0; li < CU_NCONV && ok; li++) {
        const CuLayerInfo &L = c->info[li];
        ok = need(SEC_W_F16 + li, (size_t)9 * L.cin * L.cout * 2) && need(SEC_BIAS_FUSED + li, (size_t)L.cout * 4) &&
             need(SEC_BIAS_MMA + li, (size_t)L.cout * 32);
        if (ok && L.xc > 0) ok = need(SEC_X_W_F16 + li, (size_t)L.xc * L.cout * 2 * ((li & 3) == 1 ? 2 : 1)); // shortcut weights: hi + lo
    }
    for (int i = 0; i < CU_NHEAD && ok; i++) ok = need(SEC_FC_W + i, (size_t)FC_IN[i] * FC_OUT[i] * 4) && need(SEC_FC_B + i, (size_t)FC_OUT[i] * 4);
    if (!ok) return fail(c, MLT_E_FORMAT, "'%s': missing or mis-sized section", path);
    return MLT_OK;
}

__global__ void cu_dense_descs_kernel(CtuDev *cus, const int16_t *orgpred, const int32_t *pocqp, int n, int size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CtuDev d;
    d.org = orgpred + (size_t)i * 2 * size * size;
    d.pred = d.org + (size_t)size * size;
    d.org_stride = d.pred_stride = size;
    d.poc = pocqp[2 * i];
    d.qp = pocqp[2 * i + 1];
    cus[i] = d;
}

// the whole network on n CUs stored densely at d_orgpred; results to device array `out`
int run_network(mlt_cu_ctx *c, int n, const int16_t *d_orgpred, const int32_t *d_pocqp, mlt_cu_result *out, cudaStream_t s)
{
    cu_dense_descs_kernel<<<(n + 255) / 256, 256, 0, s>>>(c->d_cus, d_orgpred, d_pocqp, n, c->size);
    CU(cudaGetLastError());
    static const bool old_stem = getenv("MLT_STEM_OLD") != nullptr; // A/B switch: the round-1 stem (conv1 and layer0.0.conv1 as two MMA stages)
    if (c->fused && old_stem)
        CU(launch_cu_stem_umma(c->size, c->d_cus, n, secp<__half>(c, SEC_STEM_CONV1), secp<__half>(c, SEC_W_F16 + 0),
                               secp<float>(c, SEC_BIAS_FUSED + 0), c->act0q, c->act[1], c->cap, c->num_sms, s));
    else if (c->fused && c->size == 16) // the composed stem with two CUs per MMA tile (stem5_cu16.cu)
        CU(launch_cu16_stem5(c->d_cus, n, secp<__half>(c, SEC_STEM5_W), secp<float>(c, SEC_STEM5_CORR), c->act0q, c->act[1], c->cap, c->num_sms, s));
    else if (c->fused) // conv1 o layer0.0.conv1 composed into one 5x5 stride-2 conv (stem5_umma.cu)
        CU(launch_cu_stem5_umma(c->size, c->d_cus, n, secp<__half>(c, SEC_STEM5_W), secp<float>(c, SEC_STEM5_CORR), secp<float>(c, SEC_STEM5_CORR) + 704,
                                c->act0q, c->act[1], c->cap, c->num_sms, s));
    else
        CU(launch_cu_conv1(c->size, c->d_cus, n, secp<__half>(c, SEC_CONV1_UMMA), c->act[0], c->cap, s));
    c->launches += 2;
    static const int stop_after = getenv("MLT_CU_STOP_AFTER") ? atoi(getenv("MLT_CU_STOP_AFTER")) : -1; // debug: run conv1 + this many convs, then synchronise
    if (stop_after >= 0) {
        const cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return fail(c, MLT_E_CUDA, "conv1 of the %d-px network (n=%d): %s", c->size, n, cudaGetErrorString(e));
    }
    for (int li = c->fused ? 1 : 0; li < CU_NCONV; li++) {
        if (stop_after >= 0 && li >= stop_after) { c->last_n = n; return MLT_OK; }
        ConvParams &p = c->conv_p[li];
        p.nimg = n;
        const cudaError_t e = launch_cu_conv(c->size, li, p, c->num_sms, s);
        if (e != cudaSuccess) return fail(c, MLT_E_CUDA, "conv %d of the %d-px network (n=%d): %s", li, c->size, n, cudaGetErrorString(e));
        if (stop_after >= 0) {
            const cudaError_t e2 = cudaStreamSynchronize(s);
            if (e2 != cudaSuccess) return fail(c, MLT_E_CUDA, "conv %d of the %d-px network (n=%d), after sync: %s", li, c->size, n, cudaGetErrorString(e2));
        }
        c->launches++;
    }
    CuHeadParams hp;
    for (int i = 0; i < CU_NHEAD; i++) {
        const int a = 4 * (i + 1) + 4; // output of layer(i+1).1.conv2 = conv 4 * (i + 1) + 3
        hp.act[i] = c->act[a];
        hp.lay[i] = c->lay[a];
        hp.hilo[i] = c->info[a - 1].hilo_out;
        hp.gap_part[i] = c->gap_part[i];
        hp.gap_count[i] = c->info[a - 1].gap_count;
        hp.fc_w[i] = secp<float>(c, SEC_FC_W + i);
        hp.fc_b[i] = secp<float>(c, SEC_FC_B + i);
    }
    hp.cus = c->d_cus; hp.out = out; hp.n = n;
    CU(launch_cu_head(hp, s));
    c->launches++;
    c->last_n = n;
    return MLT_OK;
}

int check_ctx(mlt_cu_ctx *c, bool device_entry = false)
{
    if (!c) return MLT_E_INVAL;
    c->err.clear();
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, MLT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (!device_entry && c->dev_pending_host) { // a device-resident call may still be running on the caller's stream
        for (cudaStream_t st : {c->stream, c->copy_stream}) CU(cudaStreamWaitEvent(st, c->ev_dev, 0));
        c->dev_pending_host = false;
    }
    return MLT_OK;
}

// Enqueue one dense host batch into `slot` without waiting for it: pocqp (via the slot's pinned buffer) and the chunked
// H2D copies on the copy stream, the kernels on the compute stream, the D2H of the results into the slot's pinned buffer
// and the `done` event.  `pipelined` = a batch of the submit / collect API (every copy goes to the copy stream so that it
// can run under the previous batch's kernels).
int enqueue_host_batch(mlt_cu_ctx *c, int slot_idx, int n, const int16_t *src, const int32_t *pq, bool pipelined)
{
    cudaStream_t s = c->stream;
    mlt_cu_ctx::HostSlot &H = c->slot[slot_idx];
    const size_t per = (size_t)2 * c->size * c->size;
    if (pq != H.h_pq) memcpy(H.h_pq, pq, (size_t)n * 2 * sizeof(int32_t)); // consumed before the call returns
    // Large batches are pipelined in three chunks (n/8, 3n/8, n/2): chunk i + 1 is copied (copy stream) while chunk i
    // computes, so only the first, small copy is exposed.  Few chunks on purpose: every pass over the 23 kernels has a
    // latency floor of ~0.25 ms (the last stages stream megabytes of weights through a handful of CTAs), which six
    // finer chunks paid six times (measured).  The strip buffers are reused by every chunk (their kernels are ordered
    // on the compute stream).
    int sizes[mlt_cu_ctx::MAX_CHUNKS], nchunks = 0;
    // A pipelined batch submitted while another one is still in flight needs no chunks: its whole copy hides under that
    // batch's kernels, and one pass avoids paying the latency floor three times.
    const bool hidden = pipelined && c->submitted != c->collected;
    if (hidden || (size_t)n * per * sizeof(int16_t) < ((size_t)8 << 20)) sizes[nchunks++] = n; // < 8 MiB: one copy, one pass
    else {
        sizes[0] = n / 8;
        sizes[1] = 3 * (n / 8);
        sizes[2] = n - sizes[0] - sizes[1];
        nchunks = 3;
    }
    const bool side = nchunks > 1 || pipelined;
    CU(cudaMemcpyAsync(H.d_pq, H.h_pq, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, side ? c->copy_stream : s));
    for (int i = 0, off = 0; i < nchunks; off += sizes[i], i++) {
        const int m = sizes[i];
        CU(cudaMemcpyAsync(H.d_in + (size_t)off * per, src + (size_t)off * per, (size_t)m * per * sizeof(int16_t), cudaMemcpyHostToDevice,
                           side ? c->copy_stream : s));
        if (side) {
            CU(cudaEventRecord(c->ev_in[i], c->copy_stream));
            CU(cudaStreamWaitEvent(s, c->ev_in[i], 0));
        }
        const int rc = run_network(c, m, H.d_in + (size_t)off * per, H.d_pq + (size_t)off * 2, H.d_out + off, s);
        if (rc) { cudaStreamSynchronize(c->copy_stream); return rc; }
    }
    c->last_n = nchunks == 1 ? n : 0; // debug_activation only sees a batch that ran as one pass
    CU(cudaMemcpyAsync(H.h_out, H.d_out, (size_t)n * sizeof(mlt_cu_result), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(H.done, s));
    H.n = n;
    return MLT_OK;
}

// one blocking host batch through slot 0
int run_host_batch(mlt_cu_ctx *c, int n, const int16_t *src, const int32_t *pq, mlt_cu_result *out)
{
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "%llu submitted batch(es) not collected yet", (unsigned long long)(c->submitted - c->collected));
    const int rc = enqueue_host_batch(c, 0, n, src, pq, false);
    if (rc) return rc;
    CU(cudaEventSynchronize(c->slot[0].done));
    memcpy(out, c->slot[0].h_out, (size_t)n * sizeof(mlt_cu_result));
    return MLT_OK;
}

} // namespace

extern "C" {

const char *mlt_cu_last_error(const mlt_cu_ctx *c) { return c ? c->err.c_str() : ""; }

int mlt_cu_layer_info(int cu_size, int layer, int32_t info[10])
{
    CuLayerInfo L;
    if (!info || cu_conv_info(cu_size, layer, &L) != cudaSuccess) return MLT_E_INVAL;
    const int32_t v[10] = {L.cin, L.cout, L.stride, L.hout, L.xc, L.out_par, L.nb, L.flat, L.g, L.gx};
    memcpy(info, v, sizeof v);
    return MLT_OK;
}
int mlt_cu_size(const mlt_cu_ctx *c) { return c ? c->size : 0; }
uint64_t mlt_cu_launch_count(const mlt_cu_ctx *c) { return c ? c->launches : 0; }

void mlt_cu_destroy(mlt_cu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (__half *a : c->act) cudaFree(a);
    cudaFree(c->act0q);
    for (float *g : c->gap_part) cudaFree(g);
    if (c->slot[1].d_in) {
        cudaFree(c->slot[1].d_in); cudaFree(c->slot[1].d_pq); cudaFree(c->slot[1].d_out);
        cudaFreeHost(c->slot[1].h_pq); cudaFreeHost(c->slot[1].h_out);
    }
    for (auto &H : c->slot) if (H.done) cudaEventDestroy(H.done);
    cudaFree(c->d_blob); cudaFree(c->d_in); cudaFree(c->d_pq); cudaFree(c->d_cus); cudaFree(c->d_out); cudaFree(c->d_dbg);
    cudaFree(c->d_pic); cudaFree(c->d_mv);
    cudaFreeHost(c->h_in); cudaFreeHost(c->h_pq); cudaFreeHost(c->h_out);
    for (cudaEvent_t e : c->ev_in) if (e) cudaEventDestroy(e);
    if (c->ev_dev) cudaEventDestroy(c->ev_dev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int mlt_cu_create(mlt_cu_ctx **out, const char *weights_path, int cuda_device, int cu_size, int max_batch)
{
    if (!out) return MLT_E_INVAL;
    *out = nullptr;
    if (!weights_path || (cu_size != 64 && cu_size != 32 && cu_size != 16) || max_batch < 1 || max_batch > (1 << 18)) return MLT_E_INVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MLT_E_NODEVICE;
    if (cuda_device < 0 || cuda_device >= ndev) return MLT_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cuda_device) != cudaSuccess) return MLT_E_NODEVICE;
    if (prop.major != 10) return MLT_E_NODEVICE; // sm_100a cubin only: tcgen05 / TMEM required, no other path exists
    mlt_cu_ctx *c = new (std::nothrow) mlt_cu_ctx();
    if (!c) return MLT_E_NOMEM;
    c->device = cuda_device;
    c->size = cu_size;
    c->cap = max_batch;
    c->num_sms = prop.multiProcessorCount;
    auto body = [&]() -> int {
        CU(cudaSetDevice(cuda_device));
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : c->ev_in) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_dev, cudaEventDisableTiming));
        for (int li = 0; li < CU_NCONV; li++) CU(cu_conv_info(cu_size, li, &c->info[li]));
        int r = load_blob(c, weights_path);
        if (r) return r;
        CU(conv_umma_init()); // resolves cuTensorMapEncodeTiled
        CU(cu_conv_init(cu_size));
        CU(stem_umma_init());
        CU(stem5_umma_init());
        CU(stem5_cu16_init());
        // every size runs a fused stem; the round-1 stem kernel (MLT_STEM_OLD=1) has no 16-px form (a 16-px CU is smaller than its work unit)
        c->fused = (cu_size >= 32 || getenv("MLT_STEM_OLD") == nullptr) && getenv("MLT_CU_UNFUSED") == nullptr;
        // every activation is ONE strip over the whole batch: [plane][C/8][row][cap images][x][8] fp16
        c->lay[0] = ActLayout{cu_size, 32, 1, 0, c->cap};
        for (int li = 0; li < CU_NCONV; li++) c->lay[li + 1] = ActLayout{c->info[li].hout, c->info[li].cout, c->info[li].out_par, 0, c->cap};
        for (int a = 0; a < CU_NACT; a++) {
            const bool hilo = a >= 1 && c->info[a - 1].hilo_out; // lo tensor right behind the hi tensor
            const size_t bytes = c->lay[a].unit_elems() * sizeof(__half) * (hilo ? 2 : 1);
            CU(cudaMalloc(&c->act[a], bytes));
            CU(cudaMemsetAsync(c->act[a], 0, bytes, c->stream)); // images beyond a batch's n are read by the last tile: keep them finite
        }
        c->lay0q = ActLayout{cu_size / 2, 32, 0, 0, c->cap};
        if (c->fused) {
            CU(cudaMalloc(&c->act0q, c->lay0q.unit_elems() * sizeof(__half)));
            CU(cudaMemsetAsync(c->act0q, 0, c->lay0q.unit_elems() * sizeof(__half), c->stream));
        }
        memset(c->conv_p, 0, sizeof c->conv_p);
        for (int li = 0; li < CU_NCONV; li++) {
            ConvParams &p = c->conv_p[li];
            const bool has_x = c->info[li].xc > 0; // second conv of a block: X = the block's input = activation li - 1
            // fused stem: layer0.0.conv2's shortcut reads conv1's even / even quarter instead of plane 0 of conv1's output
            const __half *xp = has_x ? ((li == 1 && c->fused) ? c->act0q : c->act[li - 1]) : nullptr;
            const ActLayout *xl = has_x ? ((li == 1 && c->fused) ? &c->lay0q : &c->lay[li - 1]) : nullptr;
            CU(cu_conv_prepare(cu_size, li, &p, c->act[li], c->lay[li], xp, xl));
            p.w = secp<__half>(c, SEC_W_F16 + li);
            p.bias = secp<__half>(c, SEC_BIAS_MMA + li);
            p.bias_f32 = secp<float>(c, SEC_BIAS_FUSED + li);
            p.x_w = has_x ? secp<__half>(c, SEC_X_W_F16 + li) : nullptr;
            p.out = c->act[li + 1];
            if (c->info[li].gap_count > 0) { // last conv of layer1..4 on a map >= 8x8: its epilogue also emits pool partial sums
                const int hi = li / 4 - 1;
                CU(cudaMalloc(&c->gap_part[hi], ((size_t)c->cap + 2) * c->info[li].gap_count * c->info[li].cout * sizeof(float)));
                p.gap_part = c->gap_part[hi];
            }
            p.relu = 1;
            p.reverse = getenv("MLT_NO_REVERSE") ? 0 : (li & 1);
            p.dbg = getenv("MLT_DEBUG_FLAGS") ? atoi(getenv("MLT_DEBUG_FLAGS")) : 0; // timing / bisect experiments only (results invalid)
        }
        const size_t per = (size_t)2 * cu_size * cu_size;
        CU(cudaMalloc(&c->d_in, (size_t)max_batch * per * sizeof(int16_t)));
        CU(cudaMalloc(&c->d_pq, (size_t)max_batch * 2 * sizeof(int32_t)));
        CU(cudaMalloc(&c->d_cus, (size_t)max_batch * sizeof(CtuDev)));
        CU(cudaMalloc(&c->d_out, (size_t)max_batch * sizeof(mlt_cu_result)));
        CU(cudaHostAlloc(&c->h_in, (size_t)max_batch * per * sizeof(int16_t), cudaHostAllocDefault));
        CU(cudaHostAlloc(&c->h_pq, (size_t)max_batch * 2 * sizeof(int32_t), cudaHostAllocDefault));
        CU(cudaHostAlloc(&c->h_out, (size_t)max_batch * sizeof(mlt_cu_result), cudaHostAllocDefault));
        c->slot[0].d_in = c->d_in; c->slot[0].d_pq = c->d_pq; c->slot[0].h_pq = c->h_pq;
        c->slot[0].d_out = c->d_out; c->slot[0].h_out = c->h_out;
        for (auto &H : c->slot) CU(cudaEventCreateWithFlags(&H.done, cudaEventDisableTiming));
        CU(cudaStreamSynchronize(c->stream));
        return MLT_OK;
    };
    const int rc = body();
    if (rc != MLT_OK) {
        fprintf(stderr, "mlt_cu_create: %s (%s)\n", mlt_strerror(rc), c->err.c_str());
        mlt_cu_destroy(c);
        return rc;
    }
    *out = c;
    return MLT_OK;
}

int mlt_cu_predict_batch(mlt_cu_ctx *c, int n, const mlt_ctu_desc *descs, mlt_cu_result *out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!descs || !out))) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->cap) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->cap);
    if (n == 0) return MLT_OK;
    const int S = c->size;
    for (int i = 0; i < n; i++) {
        if (!descs[i].org || !descs[i].pred) return fail(c, MLT_E_INVAL, "descs[%d]: null org/pred", i);
        int16_t *dst = c->h_in + (size_t)i * 2 * S * S; // private dense copy, like the hook's xMalloc buffers (EncCu.cpp:811-830)
        for (int y = 0; y < S; y++) memcpy(dst + (size_t)y * S, descs[i].org + (size_t)y * descs[i].org_stride, S * sizeof(int16_t));
        dst += (size_t)S * S;
        for (int y = 0; y < S; y++) memcpy(dst + (size_t)y * S, descs[i].pred + (size_t)y * descs[i].pred_stride, S * sizeof(int16_t));
        c->h_pq[2 * i] = descs[i].poc;
        c->h_pq[2 * i + 1] = descs[i].qp;
    }
    return run_host_batch(c, n, c->h_in, c->h_pq, out);
}

int mlt_cu_predict(mlt_cu_ctx *c, const int16_t *org, int org_stride, const int16_t *pred, int pred_stride, int poc, int qp,
                   mlt_cu_result *out)
{
    mlt_ctu_desc d;
    d.org = org; d.pred = pred; d.org_stride = org_stride; d.pred_stride = pred_stride; d.poc = poc; d.qp = qp;
    return mlt_cu_predict_batch(c, 1, &d, out);
}

int mlt_cu_predict_batch_dense(mlt_cu_ctx *c, int n, const int16_t *orgpred, const int32_t *pocqp, mlt_cu_result *out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!orgpred || !pocqp || !out))) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->cap) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->cap);
    if (n == 0) return MLT_OK;
    return run_host_batch(c, n, orgpred, pocqp, out);
}

int mlt_cu_submit_batch_dense(mlt_cu_ctx *c, int n, const int16_t *orgpred, const int32_t *pocqp)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (n <= 0 || !orgpred || !pocqp) return fail(c, MLT_E_INVAL, "bad argument");
    if (n > c->cap) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->cap);
    if (c->submitted - c->collected >= 2) return fail(c, MLT_E_STATE, "two batches are already in flight: collect one first");
    const int si = (int)(c->submitted & 1);
    mlt_cu_ctx::HostSlot &H = c->slot[si];
    if (!H.d_in) { // second slot, first use
        const size_t per = (size_t)2 * c->size * c->size;
        CU(cudaMalloc(&H.d_in, (size_t)c->cap * per * sizeof(int16_t)));
        CU(cudaMalloc(&H.d_pq, (size_t)c->cap * 2 * sizeof(int32_t)));
        CU(cudaMalloc(&H.d_out, (size_t)c->cap * sizeof(mlt_cu_result)));
        CU(cudaHostAlloc(&H.h_pq, (size_t)c->cap * 2 * sizeof(int32_t), cudaHostAllocDefault));
        CU(cudaHostAlloc(&H.h_out, (size_t)c->cap * sizeof(mlt_cu_result), cudaHostAllocDefault));
    }
    rc = enqueue_host_batch(c, si, n, orgpred, pocqp, true);
    if (rc) return rc;
    H.busy = true;
    c->submitted++;
    return MLT_OK;
}

int mlt_cu_collect(mlt_cu_ctx *c, mlt_cu_result *out, int *n_out)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!out) return fail(c, MLT_E_INVAL, "null argument");
    if (c->submitted == c->collected) return fail(c, MLT_E_STATE, "no batch in flight");
    mlt_cu_ctx::HostSlot &H = c->slot[c->collected & 1];
    CU(cudaEventSynchronize(H.done));
    memcpy(out, H.h_out, (size_t)H.n * sizeof(mlt_cu_result));
    if (n_out) *n_out = H.n;
    H.busy = false;
    c->collected++;
    return MLT_OK;
}

int mlt_cu_picture_cu_count(int cu_size, int width, int height)
{
    if ((cu_size != 64 && cu_size != 32 && cu_size != 16) || width < 0 || height < 0) return MLT_E_INVAL;
    return (width / cu_size) * (height / cu_size); // CUs of the size's raster lying fully inside the picture (EncCu.cpp:755)
}

int mlt_cu_predict_picture(mlt_cu_ctx *c, const int16_t *org_luma, int org_stride, const int16_t *ref_luma, int ref_stride, int width,
                           int height, int poc, const int16_t *mv, int qp, mlt_cu_result *out, int capacity)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!org_luma || !ref_luma || !out || width < c->size || height < c->size || org_stride < width || ref_stride < width)
        return fail(c, MLT_E_INVAL, "bad picture geometry");
    const int n = (width / c->size) * (height / c->size);
    if (n > c->cap) return fail(c, MLT_E_BATCH, "picture has %d %d-px CUs > max_batch=%d", n, c->size, c->cap);
    if (capacity < n) return fail(c, MLT_E_INVAL, "out holds %d results, the picture has %d CUs", capacity, n);
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "submitted batches must be collected first");
    cudaStream_t s = c->stream;
    const int pitch = (width + 7) & ~7; // rows 16-byte aligned on the device
    const size_t plane = (size_t)pitch * height;
    if (2 * plane > c->pic_capacity) {
        if (c->d_pic) cudaFree(c->d_pic);
        c->d_pic = nullptr;
        c->pic_capacity = 0;
        CU(cudaMalloc(&c->d_pic, 2 * plane * sizeof(int16_t)));
        c->pic_capacity = 2 * plane;
    }
    if (mv && !c->d_mv) CU(cudaMalloc(&c->d_mv, (size_t)c->cap * 2 * sizeof(int16_t)));
    int16_t *d_org = c->d_pic, *d_ref = c->d_pic + plane;
    CU(cudaMemcpy2DAsync(d_org, (size_t)pitch * sizeof(int16_t), org_luma, (size_t)org_stride * sizeof(int16_t), (size_t)width * sizeof(int16_t),
                         height, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpy2DAsync(d_ref, (size_t)pitch * sizeof(int16_t), ref_luma, (size_t)ref_stride * sizeof(int16_t), (size_t)width * sizeof(int16_t),
                         height, cudaMemcpyHostToDevice, s));
    if (mv) CU(cudaMemcpyAsync(c->d_mv, mv, (size_t)n * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    CU(launch_picture_cu_gather(d_org, d_ref, pitch, width, height, c->size, n, mv ? c->d_mv : nullptr, poc, qp, c->d_in, c->d_pq, s));
    c->launches++;
    rc = run_network(c, n, c->d_in, c->d_pq, c->d_out, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->h_out, c->d_out, (size_t)n * sizeof(mlt_cu_result), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s)); // also: the caller's planes / mv may be reused after return
    memcpy(out, c->h_out, (size_t)n * sizeof(mlt_cu_result));
    return n;
}

int mlt_cu_predict_batch_device(mlt_cu_ctx *c, int n, const int16_t *d_orgpred, const int32_t *d_pocqp, mlt_cu_result *d_out,
                                void *cuda_stream)
{
    int rc = check_ctx(c, true);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_orgpred || !d_pocqp || !d_out))) return fail(c, MLT_E_INVAL, "null argument");
    if (n > c->cap) return fail(c, MLT_E_BATCH, "n=%d > max_batch=%d", n, c->cap);
    if (((uintptr_t)d_orgpred & 15) != 0) return fail(c, MLT_E_INVAL, "d_orgpred must be 16-byte aligned");
    if (c->submitted != c->collected) return fail(c, MLT_E_STATE, "submitted batches must be collected first");
    if (n == 0) return MLT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (c->dev_issued) CU(cudaStreamWaitEvent(s, c->ev_dev, 0)); // after the previous device call, whatever its stream
    rc = run_network(c, n, d_orgpred, d_pocqp, d_out, s);
    if (cudaEventRecord(c->ev_dev, s) == cudaSuccess) c->dev_issued = c->dev_pending_host = true;
    return rc;
}

int64_t mlt_cu_debug_activation(mlt_cu_ctx *c, int layer, float *out, int64_t capacity)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (layer < 0 || layer >= CU_NACT || !out) return fail(c, MLT_E_INVAL, "bad layer");
    if (c->last_n <= 0) return fail(c, MLT_E_STATE, "no batch has been run");
    const ActLayout &L = c->lay[layer];
    const size_t elems = (size_t)L.H * L.H * L.C * c->last_n;
    if ((int64_t)elems > capacity) return fail(c, MLT_E_INVAL, "capacity %lld < %zu", (long long)capacity, elems);
    if (c->dbg_bytes < elems * sizeof(float)) {
        if (c->d_dbg) cudaFree(c->d_dbg);
        c->d_dbg = nullptr;
        c->dbg_bytes = 0;
        CU(cudaMalloc(&c->d_dbg, elems * sizeof(float)));
        c->dbg_bytes = elems * sizeof(float);
    }
    cudaStream_t s = c->stream;
    if (layer == 0 && c->fused) // the fused stem never materialises conv1's output: recompute it with the standalone kernel
        CU(launch_cu_conv1(c->size, c->d_cus, c->last_n, secp<__half>(c, SEC_CONV1_UMMA), c->act[0], c->cap, s));
    CU(launch_unpack_act(c->act[layer], c->d_dbg, c->last_n, L, s));
    CU(cudaMemcpyAsync(out, c->d_dbg, elems * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return (int64_t)elems;
}

} // extern "C"
