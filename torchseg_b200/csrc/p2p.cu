// Device-initiated all-reduce of SMALL fp32 vectors over NVLink peer memory — the SyncBatchNorm statistics exchange
// (replaces apex SyncBatchNorm's per-layer all-reduce, /root/reference/model/bisenet/cityscapes.bisenet.R18/train.py:54-55).
//
// A BiSeNet step has ~70 of these exchanges ([Σx | Σx²] forward, [Σdz | Σdz·x̂] backward; 128..8192 floats each): with NCCL
// every one costs a launch + protocol latency of 25-40 us on the compute stream, i.e. the whole weak-scaling loss of the
// step. Here ONE small kernel per exchange writes the local partial sums straight into every peer's symmetric buffer
// (NVLink stores), publishes a per-(slot, writer) sequence flag with release semantics, spins on the flags of all peers
// and reduces the `world` partials in rank order (identical result bits on every rank, run-to-run deterministic).
//
// Symmetric buffer of every rank (allocated and exchanged by the caller, e.g. torch symmetric memory):
//   [nslots] x { float data[world][slot_floats];  uint32 flag[world] (padded to 128 B) }
// Exchange number `seq` (1, 2, 3, ... identical on all ranks: every rank executes the same layer sequence) uses slot
// seq % nslots and flag value seq, so flags never need resetting; a slot is re-used only after nslots-1 other exchanges,
// each of which orders all ranks, so a writer can never overtake a reader of the previous use.
#include "tsb_common.cuh"

namespace {

constexpr int kXThreads = 1024;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

struct XchgParams {
    float* vals;                 // [n] local partial sums in, global sums out (n % 4 == 0)
    int n;
    unsigned long long peers[TSB_P2P_MAX_WORLD];   // symmetric buffer base address on every rank (peer-mapped)
    int rank, world;
    unsigned int seq;
    int nslots, slot_floats;
    float* acc_hi;               // optional: acc_hi[c] += vals[half + c]  (LOCAL Σdz·x̂ → dgamma) before the exchange
    float* acc_lo;               //           acc_lo[c] += vals[c]         (LOCAL Σdz   → dbeta)
    int half;
};

__global__ void __launch_bounds__(kXThreads, 1) p2p_allreduce_kernel(const XchgParams p) {
    const int tid = threadIdx.x;
    const size_t slot_bytes = ((size_t)p.world * p.slot_floats * 4 + (size_t)p.world * 4 + 127) / 128 * 128;
    const size_t slot_off = (size_t)(p.seq % (unsigned)p.nslots) * slot_bytes;
    const size_t flag_off = slot_off + (size_t)p.world * p.slot_floats * 4;
    const int n4 = p.n >> 2;
    // 1. local parameter-gradient accumulation (local sums: the DDP all-reduce averages parameter gradients later)
    if (p.acc_hi != nullptr)
        for (int c = tid; c < p.half; c += kXThreads) {
            p.acc_hi[c] += p.vals[p.half + c];
            p.acc_lo[c] += p.vals[c];
        }
    // 2. my partial → region [rank] of the slot on every rank (own copy included)
    for (int i = tid; i < n4; i += kXThreads) {
        const float4 v = *reinterpret_cast<const float4*>(p.vals + 4 * i);
        for (int r = 0; r < p.world; ++r) {
            float* dst = reinterpret_cast<float*>(p.peers[r] + slot_off) + (size_t)p.rank * p.slot_floats + 4 * i;
            *reinterpret_cast<float4*>(dst) = v;
        }
    }
    __syncthreads();
    // 3. publish: one release store per peer (orders the CTA's data stores, made visible system-wide by the fence)
    if (tid < p.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(p.peers[tid] + flag_off) + p.rank, p.seq);
    }
    // 4. wait for every writer's flag in MY buffer (bounded: a dead peer traps instead of hanging the box)
    if (tid < p.world) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(p.peers[p.rank] + flag_off) + tid;
        unsigned long long t0 = 0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(f) != p.seq) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 30ull * 1000000000ull) __trap();   // 30 s: a peer died or the ranks diverged
        }
    }
    __syncthreads();
    // 5. reduce in rank order (L1-bypassing loads: the lines were written by peers)
    const float* mine = reinterpret_cast<const float*>(p.peers[p.rank] + slot_off);
    for (int i = tid; i < n4; i += kXThreads) {
        float4 s = ld_relaxed_sys_v4(mine + 4 * i);
        for (int r = 1; r < p.world; ++r) {
            const float4 v = ld_relaxed_sys_v4(mine + (size_t)r * p.slot_floats + 4 * i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4*>(p.vals + 4 * i) = s;
    }
}

}  // namespace

extern "C" size_t tsb_p2p_buffer_bytes(int world, int nslots, int slot_floats) {
    if (world < 1 || nslots < 2 || slot_floats < 4) return 0;
    const size_t slot_bytes = ((size_t)world * slot_floats * 4 + (size_t)world * 4 + 127) / 128 * 128;
    return slot_bytes * (size_t)nslots;
}

extern "C" int tsb_p2p_allreduce_sum(float* vals, int n, const unsigned long long* peer_bases, int rank, int world,
                                     unsigned int seq, int nslots, int slot_floats, float* acc_hi, float* acc_lo,
                                     tsb_stream_t stream) {
    TSB_REQUIRE(vals && peer_bases, "tsb_p2p_allreduce_sum: null pointer");
    TSB_REQUIRE(world >= 1 && world <= TSB_P2P_MAX_WORLD && rank >= 0 && rank < world, "tsb_p2p_allreduce_sum: bad rank/world");
    TSB_REQUIRE(n > 0 && n % 4 == 0 && n <= slot_floats && slot_floats % 4 == 0, "tsb_p2p_allreduce_sum: n must be a multiple of 4 and fit a slot (n=%d, slot=%d)", n, slot_floats);
    TSB_REQUIRE(nslots >= 2 && seq != 0, "tsb_p2p_allreduce_sum: nslots >= 2, seq >= 1");
    TSB_REQUIRE((acc_hi == nullptr) == (acc_lo == nullptr) && (acc_hi == nullptr || n % 2 == 0), "tsb_p2p_allreduce_sum: accumulators go together");
    TSB_REQUIRE(tsb_aligned16(vals), "tsb_p2p_allreduce_sum: vals must be 16-byte aligned");
    XchgParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.vals = vals; prm.n = n;
    for (int r = 0; r < world; ++r) {
        TSB_REQUIRE(peer_bases[r] != 0 && peer_bases[r] % 128 == 0, "tsb_p2p_allreduce_sum: peer base %d missing / not 128-byte aligned", r);
        prm.peers[r] = peer_bases[r];
    }
    prm.rank = rank; prm.world = world; prm.seq = seq; prm.nslots = nslots; prm.slot_floats = slot_floats;
    prm.acc_hi = acc_hi; prm.acc_lo = acc_lo; prm.half = n / 2;
    p2p_allreduce_kernel<<<1, kXThreads, 0, (cudaStream_t)stream>>>(prm);
    TSB_CUDA_CHECK_LAUNCH("p2p_allreduce");
    return TSB_OK;
}
