// Exchange step over NVLink / NVSwitch peer memory: a one-shot all-reduce written as one kernel.
//
// The payload of the path's single exchange is small (mean-field: 4 D + 8 floats = 16 KB at
// D = 1025; full-rank: a few MB), i.e. latency-bound, and it sits in the middle of a captured
// CUDA-graph iteration.  Every rank owns a "symmetric" buffer (two data slots + a flag word per
// peer) that all peers map through CUDA IPC.  One launch per exchange:
//   1. copy the local partial sums into my slot (seq & 1),
//   2. the last CTA to finish publishes: a release store of `seq` into flag[my_rank] of EVERY peer,
//   3. every CTA waits until all flags in my own array reached `seq` (acquire loads, local memory),
//   4. every rank sums the nranks slots over NVLink in rank order 0..n-1  => bit-identical result
//      on all ranks, independent of arrival order (keeps "same seed => same run").
// Slot parity replaces a second barrier: a rank can only overwrite slot s at seq+2 after every
// peer published seq+1, which they do after finishing their reads of seq.
// No NCCL call, no host involvement: safe inside a graph.
#include <cstring>

#include "avi_internal.cuh"
#include "comm_dev.cuh"

namespace {

constexpr int MAX_RANKS = AVI_MAX_RANKS;
constexpr int FLAG_WORDS = AVI_FLAG_WORDS;

struct CommState {
    int rank = 0, nranks = 1;
    int64_t max_floats = 0;
    int64_t ll_cap = 0;          // floats per LL lane
    void* base = nullptr;        // my symmetric allocation: [flags (256 B) | slot 0 | slot 1 | LL area]
    void* peer_base[MAX_RANKS] = {nullptr};
    bool opened[MAX_RANKS] = {false};
    CommDev* dev = nullptr;
    float* bar_buf = nullptr;    // payload of avi_comm_barrier
    PeerTable table{};
    bool connected = false;
};

// Small payloads: the low-latency push protocol (comm_dev.cuh).  Every thread pushes its elements to all peers, then
// sums the ranks' values in rank order as they arrive; the last CTA to finish advances the sequence number.
__global__ void __launch_bounds__(256)
k_allreduce_ll(float* __restrict__ buf, long long count, CommPeers c) {
    const unsigned int seq = *reinterpret_cast<volatile unsigned int*>(&c.dev->seq) + 1u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = i0; i < count; i += stride) ll_push(c, seq, i, buf[i]);
    for (long long i = i0; i < count; i += stride) {
        const long long idx[1] = {i}; const bool need[1] = {true}; const float own[1] = {buf[i]};
        float out[1];
        ll_gather<1>(c, seq, idx, need, own, out);
        buf[i] = out[0];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&c.dev->depart, 1u) == gridDim.x - 1) {
            c.dev->depart = 0;
            *reinterpret_cast<volatile unsigned int*>(&c.dev->seq) = seq;
        }
    }
}

// Larger payloads (full-rank family: the M x D gradient block or the D x D contraction, a few MB): pull protocol.
// Data moves as 16-byte words and a thread requests the slots of ALL ranks for its word before it adds them up (in
// rank order), so one NVLink round trip covers the whole payload: the first version walked 4-byte words with a
// dependent add after every load (16 serialised round trips per thread: +60 us per step on C3 at 2 GPUs).
// The grid must be co-resident (every CTA copies before anyone's flag wait can end): <= 2 CTAs per SM.
__global__ void __launch_bounds__(256)
k_allreduce_oneshot(float* __restrict__ buf, long long count, long long slot_stride, PeerTable t, int rank,
                    int nranks, CommDev* __restrict__ cd) {
    const unsigned int seq = *reinterpret_cast<volatile unsigned int*>(&cd->seq) + 1u;
    const long long off = (long long)(seq & 1u) * slot_stride;
    float* mine = t.data[rank] + off;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = (reinterpret_cast<uintptr_t>(buf) & 15) == 0;   // (slots are 256-byte aligned by construction)
    const long long n4 = vec ? count / 4 : 0;
    for (long long i = tid; i < n4; i += stride)
        reinterpret_cast<float4*>(mine)[i] = reinterpret_cast<const float4*>(buf)[i];
    for (long long i = 4 * n4 + tid; i < count; i += stride) mine[i] = buf[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&cd->arrive, 1u) == gridDim.x - 1) {
            cd->arrive = 0;
            __threadfence_system();
            for (int r = 0; r < nranks; ++r) st_release_sys(t.flags[r] + rank, seq);
        }
    }
    if (threadIdx.x < nranks) {
        const unsigned int* f = t.flags[rank] + threadIdx.x;
        while ((int)(ld_acquire_sys(f) - seq) < 0) { }
    }
    __syncthreads();
    for (long long i = tid; i < n4; i += stride) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r0 = 0; r0 < MAX_RANKS; r0 += 8) {
            if (r0 >= nranks) break;
            float4 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < nranks) v[r] = ld_relaxed_sys_v4(t.data[r0 + r] + off + 4 * i);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < nranks) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        }
        reinterpret_cast<float4*>(buf)[i] = s;
    }
    for (long long i = 4 * n4 + tid; i < count; i += stride) {
        float s = 0.0f;
        for (int r = 0; r < nranks; ++r) s += ld_relaxed_sys(t.data[r] + off + i);
        buf[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&cd->depart, 1u) == gridDim.x - 1) {
            cd->depart = 0;
            *reinterpret_cast<volatile unsigned int*>(&cd->seq) = seq;
        }
    }
}

CommState* state(avi_ctx* ctx) { return static_cast<CommState*>(ctx->comm); }

int32_t finish_connect(avi_ctx* ctx, CommState* cs) {
    for (int r = 0; r < cs->nranks; ++r) {
        cs->table.flags[r] = static_cast<unsigned int*>(cs->peer_base[r]);
        cs->table.data[r] = reinterpret_cast<float*>(static_cast<char*>(cs->peer_base[r]) + FLAG_WORDS * 4);
        cs->table.ll[r] = reinterpret_cast<LLWord*>(static_cast<char*>(cs->peer_base[r]) + FLAG_WORDS * 4 +
                                                    2 * (size_t)cs->max_floats * sizeof(float));
    }
    cs->connected = true;
    ctx->rank = cs->rank; ctx->nranks = cs->nranks;
    ctx->comm_capturable = true;
    return AVI_OK;
}

}  // namespace

// peers for a kernel that fuses the exchange (payload of up to max_floats floats); false when not connected
bool avi_comm_peers(avi_ctx* ctx, int64_t count, CommPeers* out) {
    CommState* cs = state(ctx);
    if (!cs || !cs->connected || count > cs->max_floats) return false;
    out->nranks = cs->nranks; out->rank = cs->rank; out->slot_stride = cs->max_floats; out->t = cs->table; out->dev = cs->dev;
    static const bool ll_off = getenv("AVI_COMM_LL") && atoi(getenv("AVI_COMM_LL")) == 0;
    out->ll_cap = ll_off ? 0 : cs->ll_cap;
    return true;
}

// largest payload (floats) the native exchange takes; -1 when the peers are not connected
int64_t avi_comm_capacity(avi_ctx* ctx) {
    CommState* cs = state(ctx);
    return cs && cs->connected ? cs->max_floats : -1;
}

int32_t avi_comm_exchange(avi_ctx* ctx, float* buf, int64_t count) {
    CommState* cs = state(ctx);
    if (!cs || !cs->connected) return AVI_ERR_UNSUPPORTED;
    if (count > cs->max_floats) AVI_FAIL(ctx, AVI_ERR_COMM, "exchange payload larger than the symmetric buffer");
    int grid = (int)std::min<int64_t>(ceil_div(count, 256 * 4), 2 * (int64_t)ctx->prop.multiProcessorCount);
    if (grid < 1) grid = 1;
    CommPeers peers;
    if (avi_comm_peers(ctx, count, &peers) && peers.ll_cap >= count) {
        k_allreduce_ll<<<(int)std::min<int64_t>(ceil_div(count, 256), 32), 256, 0, ctx->stream>>>(buf, count, peers);
        AVI_LAUNCHED(ctx);
        return AVI_OK;
    }
    k_allreduce_oneshot<<<grid, 256, 0, ctx->stream>>>(buf, count, cs->max_floats, cs->table, cs->rank, cs->nranks, cs->dev);
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

void avi_comm_destroy(avi_ctx* ctx) {
    CommState* cs = state(ctx);
    if (!cs) return;
    for (int r = 0; r < MAX_RANKS; ++r)
        if (cs->opened[r]) cudaIpcCloseMemHandle(cs->peer_base[r]);
    if (cs->base) cudaFree(cs->base);
    if (cs->dev) cudaFree(cs->dev);
    if (cs->bar_buf) cudaFree(cs->bar_buf);
    delete cs;
    ctx->comm = nullptr;
}

extern "C" {

// Allocate this rank's symmetric buffer for payloads of up to max_floats floats and export its
// CUDA IPC handle (64 bytes) for the peers.
int32_t avi_comm_buffer(avi_ctx* ctx, int64_t max_floats, char* handle_out) {
    if (!ctx || max_floats <= 0 || !handle_out) return AVI_ERR_INVALID;
    cudaSetDevice(ctx->device);
    avi_comm_destroy(ctx);
    CommState* cs = new CommState();
    ctx->comm = cs;
    cs->max_floats = round_up(max_floats, 64);
    cs->ll_cap = std::min<int64_t>(cs->max_floats, 16384);   // LL lanes for payloads of up to 64 KB
    const size_t ll_bytes = 2 * (size_t)MAX_RANKS * (size_t)cs->ll_cap * sizeof(LLWord);
    const size_t bytes = FLAG_WORDS * 4 + 2 * (size_t)cs->max_floats * sizeof(float) + ll_bytes;
    AVI_CHECK(avi_dev_alloc(ctx, &cs->base, bytes));
    // avi_dev_alloc zero-fills on the stream (sequence numbers start at 1: zero = nothing yet); the fill must have
    // happened before any peer can push into the buffer, i.e. before the handle leaves this call
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    AVI_CHECK(avi_alloc(ctx, &cs->dev, 1));
    AVI_CHECK(avi_alloc(ctx, &cs->bar_buf, 16));
    cudaIpcMemHandle_t h;
    AVI_CUDA(ctx, cudaIpcGetMemHandle(&h, cs->base));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    std::memcpy(handle_out, &h, 64);
    return AVI_OK;
}

// Device-side rendezvous of the ranks on the ctx stream: a one-float exchange (every rank leaves it within an NVLink
// store latency of the last rank entering).  Measurement harnesses use it to start a timed region together.
int32_t avi_comm_barrier(avi_ctx* ctx) {
    if (!ctx) return AVI_ERR_INVALID;
    CommState* cs = state(ctx);
    if (!cs || !cs->connected || cs->nranks <= 1) return AVI_OK;
    cudaSetDevice(ctx->device);
    return avi_comm_exchange(ctx, cs->bar_buf, 1);
}

// Unmap every peer's buffer (this rank's own allocation stays).  Ranks call this -- and then synchronise among
// themselves -- before any of them frees or re-creates its buffer, so that no exporter frees memory a peer still maps.
int32_t avi_comm_disconnect(avi_ctx* ctx) {
    if (!ctx) return AVI_ERR_INVALID;
    CommState* cs = state(ctx);
    if (!cs) return AVI_OK;
    cudaSetDevice(ctx->device);
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < MAX_RANKS; ++r)
        if (cs->opened[r]) { cudaIpcCloseMemHandle(cs->peer_base[r]); cs->opened[r] = false; cs->peer_base[r] = nullptr; }
    cs->connected = false;
    ctx->comm_capturable = false;
    ctx->rank = 0; ctx->nranks = 1;
    return AVI_OK;
}

// handles: nranks x 64 bytes, entry r exported by rank r (all-gathered by the host: torch.distributed
// in the Python mirror, MPI.jl / Distributed from Julia).  Entry `rank` is this rank's own.
int32_t avi_comm_connect(avi_ctx* ctx, int32_t rank, int32_t nranks, const char* handles) {
    if (!ctx || !handles || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) return AVI_ERR_INVALID;
    CommState* cs = state(ctx);
    if (!cs || !cs->base) AVI_FAIL(ctx, AVI_ERR_STATE, "call avi_comm_buffer first");
    cudaSetDevice(ctx->device);
    cs->rank = rank; cs->nranks = nranks;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { cs->peer_base[r] = cs->base; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * (size_t)r, 64);
        AVI_CUDA(ctx, cudaIpcOpenMemHandle(&cs->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
        cs->opened[r] = true;
    }
    return finish_connect(ctx, cs);
}

}  // extern "C"
