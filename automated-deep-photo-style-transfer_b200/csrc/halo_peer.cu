// Halo exchange of column strips over peer memory (NVLink / NVSwitch), for the spatially tiled optimisation (tiled.py).
//
// The reference has no multi-device path; this is the exchange step that the column-strip decomposition of
// its per-iteration step (style_transfer.py:331-344) needs between two segments of the network.
//
// Every rank owns one MAILBOX: a cudaMalloc'ed buffer that its two neighbours map through CUDA IPC (other processes) or use
// directly (ranks of one process).  Layout: [flag words][data of the left neighbour][data of the right neighbour].
//   push  one kernel on the sender: the `hl` own columns next to each interior boundary are stored straight into the
//         neighbour's mailbox (remote stores), and when the last block of a side has finished, its sequence number is released
//         into the neighbour's flag word for that (side, slot).
//   pull  one kernel on the receiver: waits until the flag of (side, slot) has reached the receiver's own sequence number, copies
//         the slab from the local mailbox into the halo columns of the tensor and folds max|slab| into the tensor's scale slot
//         (what adpst_absmax_update does after an NCCL exchange).
// Sequence numbers live in device memory on both sides (sent[] on the sender, taken[] on the receiver, both advanced by the last
// block of the kernel), so a captured CUDA graph can be replayed: no kernel argument changes from step to step.
//
// A slot is single-buffered.  That is safe when every step makes at least two exchanges with each neighbour and the ranks run the
// same sequence of exchanges: sender A's push(k, step t+1) is stream-ordered after a pull of A that waited for a push of B which
// B enqueued after its pull(k, step t).
#include <cstring>

#include "adpst.h"
#include "common.cuh"

namespace adpst {

constexpr int HALO_FLAG_BYTES = 4096;
constexpr int HALO_MAX_SLOTS = 64;                   // per side
// flag words inside the first HALO_FLAG_BYTES of the mailbox (all uint32):
//   arrived[side][slot]   written by the neighbour on that side (remote)
// local (never touched by a peer), in a second allocation: sent[side][slot], taken[side][slot], ticket[2][side][slot]

struct HaloPeer {
    char* remote = nullptr;                          // mapped mailbox of the neighbour (nullptr: no neighbour on that side)
    bool ipc = false;
};

}  // namespace adpst

struct adpst_halo {
    int device = 0;
    size_t side_bytes = 0;                           // capacity of one side's data region
    char* mailbox = nullptr;                         // HALO_FLAG_BYTES + 2 * side_bytes
    uint32_t* local = nullptr;                       // sent[2][S], taken[2][S], push_ticket[2][S], pull_ticket[2][S]
    adpst::HaloPeer peer[2];                         // 0 = left neighbour, 1 = right neighbour
};

namespace adpst {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t halo_timer_hi() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%globaltimer_hi;" : "=r"(v));
    return v;
}

struct HaloSide {
    char* remote_data;          // push: slab region in the neighbour's mailbox   | pull: slab region in the local mailbox
    uint32_t* flag;             // push: neighbour's arrived[side'][slot]          | pull: local arrived[side][slot]
    uint32_t* seq;              // push: local sent[side][slot]                    | pull: local taken[side][slot]
    uint32_t* ticket;           // local block counter
    int col0;                   // first column of the slab in the tensor
};

// tensor: [rows][width][C] float32; slab: [rows][hl * C] packed.  One float4 per thread and step; gridDim.y = side.
__global__ void __launch_bounds__(256) halo_push_kernel(const float* __restrict__ x, HaloSide s0, HaloSide s1, int rows, int width,
                                                        int C, int hl) {
    const HaloSide s = blockIdx.y == 0 ? s0 : s1;
    if (s.remote_data == nullptr) return;
    const int row4 = hl * C / 4;                                      // float4 per slab row
    const size_t total = size_t(rows) * row4;
    float4* dst = reinterpret_cast<float4*>(s.remote_data);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t r = i / row4;
        const int j = int(i - r * row4);
        dst[i] = __ldg(reinterpret_cast<const float4*>(x + (r * width + s.col0) * C) + j);
    }
    // the last block of this side publishes the slab
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(s.ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            const uint32_t seq = *s.seq + 1u;
            *s.seq = seq;
            *s.ticket = 0u;
            st_release_sys(s.flag, seq);
        }
    }
}

__global__ void __launch_bounds__(256) halo_pull_kernel(float* __restrict__ x, HaloSide s0, HaloSide s1, int rows, int width, int C,
                                                        int hl, uint32_t* __restrict__ absmax_slot) {
    const HaloSide s = blockIdx.y == 0 ? s0 : s1;
    if (s.remote_data == nullptr) return;
    if (threadIdx.x == 0) {
        const uint32_t want = *s.seq + 1u;                            // *s.seq is advanced only after every block has passed
        const uint32_t t0 = halo_timer_hi();
        while (int32_t(ld_acquire_sys(s.flag) - want) < 0) {
            if (halo_timer_hi() - t0 > 4u) __trap();                  // > ~17 s: the neighbour is gone
        }
    }
    __syncthreads();
    const int row4 = hl * C / 4;
    const size_t total = size_t(rows) * row4;
    const float4* src = reinterpret_cast<const float4*>(s.remote_data);
    float m = 0.f;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t r = i / row4;
        const int j = int(i - r * row4);
        const float4 v = __ldcv(src + i);                             // written by a peer: never from a stale L1 line
        reinterpret_cast<float4*>(x + (r * width + s.col0) * C)[j] = v;
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (absmax_slot != nullptr) {
        const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        if ((threadIdx.x & 31) == 0 && wm != 0u) atomicMax(absmax_slot, wm);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(s.ticket, 1u);
        if (t == gridDim.x - 1) {
            *s.seq = *s.seq + 1u;
            *s.ticket = 0u;
        }
    }
}

static inline uint32_t* arrived_of(char* mailbox, int side, int slot) {
    return reinterpret_cast<uint32_t*>(mailbox) + side * HALO_MAX_SLOTS + slot;
}
static inline char* data_of(char* mailbox, size_t side_bytes, int side, size_t offset) {
    return mailbox + HALO_FLAG_BYTES + size_t(side) * side_bytes + offset;
}

}  // namespace adpst

extern "C" {

int adpst_halo_create(size_t side_bytes, adpst_halo** out) {
    using namespace adpst;
    ADPST_REQUIRE(out && side_bytes > 0 && side_bytes % 16 == 0, "halo_create: bad argument");
    static_assert(2 * HALO_MAX_SLOTS * sizeof(uint32_t) <= HALO_FLAG_BYTES, "flag area too small");
    adpst_halo* h = new adpst_halo();
    h->device = current_device();
    h->side_bytes = side_bytes;
    // plain cudaMalloc: memory of a stream-ordered pool cannot be exported through cudaIpcGetMemHandle
    ADPST_CUDA_CHECK(cudaMalloc(&h->mailbox, HALO_FLAG_BYTES + 2 * side_bytes));
    ADPST_CUDA_CHECK(cudaMemset(h->mailbox, 0, HALO_FLAG_BYTES));
    ADPST_CUDA_CHECK(cudaMalloc(&h->local, 4 * 2 * HALO_MAX_SLOTS * sizeof(uint32_t)));
    ADPST_CUDA_CHECK(cudaMemset(h->local, 0, 4 * 2 * HALO_MAX_SLOTS * sizeof(uint32_t)));
    ADPST_CUDA_CHECK(cudaDeviceSynchronize());
    *out = h;
    return ADPST_OK;
}

void adpst_halo_destroy(adpst_halo* h) {
    if (!h) return;
    for (int s = 0; s < 2; ++s)
        if (h->peer[s].remote && h->peer[s].ipc) cudaIpcCloseMemHandle(h->peer[s].remote);
    cudaFree(h->mailbox);
    cudaFree(h->local);
    delete h;
}

int adpst_halo_ipc_handle_bytes(void) { return int(sizeof(cudaIpcMemHandle_t)); }

int adpst_halo_export(const adpst_halo* h, void* handle_out) {
    using namespace adpst;
    ADPST_REQUIRE(h && handle_out, "halo_export: NULL argument");
    cudaIpcMemHandle_t ipc;
    ADPST_CUDA_CHECK(cudaIpcGetMemHandle(&ipc, h->mailbox));
    std::memcpy(handle_out, &ipc, sizeof(ipc));
    return ADPST_OK;
}

int adpst_halo_connect_ipc(adpst_halo* h, int side, const void* handle, size_t peer_side_bytes) {
    using namespace adpst;
    ADPST_REQUIRE(h && handle && (side == 0 || side == 1), "halo_connect_ipc: bad argument");
    ADPST_REQUIRE(peer_side_bytes == h->side_bytes, "halo_connect_ipc: the neighbour's mailbox has another capacity");
    ADPST_REQUIRE(h->peer[side].remote == nullptr, "halo_connect_ipc: side %d is connected already", side);
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, handle, sizeof(ipc));
    void* p = nullptr;
    ADPST_CUDA_CHECK(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->peer[side].remote = static_cast<char*>(p);
    h->peer[side].ipc = true;
    return ADPST_OK;
}

int adpst_halo_connect_local(adpst_halo* h, int side, adpst_halo* neighbour) {
    using namespace adpst;
    ADPST_REQUIRE(h && neighbour && (side == 0 || side == 1), "halo_connect_local: bad argument");
    ADPST_REQUIRE(neighbour->side_bytes == h->side_bytes, "halo_connect_local: the neighbour's mailbox has another capacity");
    ADPST_REQUIRE(h->peer[side].remote == nullptr, "halo_connect_local: side %d is connected already", side);
    if (neighbour->device != h->device) {
        int can = 0;
        ADPST_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, h->device, neighbour->device));
        ADPST_REQUIRE(can, "halo_connect_local: device %d cannot reach device %d", h->device, neighbour->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(neighbour->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
        else ADPST_CUDA_CHECK(e);
    }
    h->peer[side].remote = neighbour->mailbox;
    h->peer[side].ipc = false;
    return ADPST_OK;
}

static int halo_check(const adpst_halo* h, int slot, size_t offset, int rows, int width, int C, int hl, const char* who) {
    using namespace adpst;
    ADPST_REQUIRE(h, "%s: NULL handle", who);
    ADPST_REQUIRE(slot >= 0 && slot < HALO_MAX_SLOTS, "%s: slot %d out of range", who, slot);
    ADPST_REQUIRE(rows > 0 && width > 0 && C > 0 && hl > 0 && (hl * C) % 4 == 0 && 2 * hl <= width, "%s: bad slab %dx%dx%d in width %d",
                  who, rows, hl, C, width);
    ADPST_REQUIRE(offset % 16 == 0 && offset + size_t(rows) * hl * C * sizeof(float) <= h->side_bytes,
                  "%s: slab of %zu bytes at offset %zu exceeds the mailbox side of %zu bytes", who,
                  size_t(rows) * hl * C * sizeof(float), offset, h->side_bytes);
    return ADPST_OK;
}

static unsigned halo_blocks(int rows, int hl, int C) {
    const size_t want = (size_t(rows) * hl * C / 4 + 255) / 256;
    const size_t cap = size_t(adpst::num_sms());                      // a slab is a few MB at most: one block per SM and side
    return unsigned(want < cap ? (want ? want : 1) : cap);
}

int adpst_halo_push(adpst_halo* h, int slot, size_t offset, const float* x_dev, int rows, int width, int C, int hl,
                    int own_lo, int own_hi, adpst_stream_t stream) {
    using namespace adpst;
    int rc = halo_check(h, slot, offset, rows, width, C, hl, "halo_push");
    if (rc != ADPST_OK) return rc;
    ADPST_REQUIRE(x_dev && own_lo >= 0 && own_lo + hl <= own_hi && own_hi <= width, "halo_push: bad own columns %d..%d", own_lo, own_hi);
    HaloSide s[2];
    for (int side = 0; side < 2; ++side) {
        // what goes to the left neighbour arrives there "from the right" (side 1 of its mailbox) and vice versa
        const bool on = h->peer[side].remote != nullptr && (side == 0 ? own_lo > 0 : own_hi < width);
        s[side].remote_data = on ? data_of(h->peer[side].remote, h->side_bytes, 1 - side, offset) : nullptr;
        s[side].flag = on ? arrived_of(h->peer[side].remote, 1 - side, slot) : nullptr;
        s[side].seq = h->local + (0 * 2 + side) * HALO_MAX_SLOTS + slot;
        s[side].ticket = h->local + (2 * 2 + side) * HALO_MAX_SLOTS + slot;
        s[side].col0 = side == 0 ? own_lo : own_hi - hl;
    }
    if (!s[0].remote_data && !s[1].remote_data) return ADPST_OK;
    halo_push_kernel<<<dim3(halo_blocks(rows, hl, C), 2), 256, 0, as_stream(stream)>>>(x_dev, s[0], s[1], rows, width, C, hl);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_halo_pull(adpst_halo* h, int slot, size_t offset, float* x_dev, int rows, int width, int C, int hl, int own_lo,
                    int own_hi, uint32_t* absmax_slot_dev, adpst_stream_t stream) {
    using namespace adpst;
    int rc = halo_check(h, slot, offset, rows, width, C, hl, "halo_pull");
    if (rc != ADPST_OK) return rc;
    ADPST_REQUIRE(x_dev && own_lo >= 0 && own_lo <= own_hi && own_hi <= width, "halo_pull: bad own columns %d..%d", own_lo, own_hi);
    HaloSide s[2];
    for (int side = 0; side < 2; ++side) {
        const bool on = h->peer[side].remote != nullptr && (side == 0 ? own_lo >= hl : own_hi + hl <= width) &&
                        (side == 0 ? own_lo > 0 : own_hi < width);
        s[side].remote_data = on ? data_of(h->mailbox, h->side_bytes, side, offset) : nullptr;
        s[side].flag = arrived_of(h->mailbox, side, slot);
        s[side].seq = h->local + (1 * 2 + side) * HALO_MAX_SLOTS + slot;
        s[side].ticket = h->local + (3 * 2 + side) * HALO_MAX_SLOTS + slot;
        s[side].col0 = side == 0 ? own_lo - hl : own_hi;
    }
    if (!s[0].remote_data && !s[1].remote_data) return ADPST_OK;
    halo_pull_kernel<<<dim3(halo_blocks(rows, hl, C), 2), 256, 0, as_stream(stream)>>>(x_dev, s[0], s[1], rows, width, C, hl,
                                                                                     absmax_slot_dev);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

}  // extern "C"
