/*
 * exchange.cuh -- kernel K6: the one exchange step this path can have.
 *
 * ONE long capture is split in time across the GPUs of a box; every GPU runs k_spectrum over the
 * frames of its own time range, and the 1024 bin sums have to be added across GPUs.  Instead of
 * k_spectrum_finalize followed by a library all-reduce, this kernel does both: it reduces the local
 * per-CTA partials (fixed order), PUSHES the 1024 scaled sums straight into a mailbox slot in every
 * peer's HBM with ordinary stores over NVLink (peer memory mapped through CUDA IPC, one process per
 * GPU), raises a flag in every peer, waits for the flags of all ranks in its own mailbox and adds
 * the slots in rank order.  Every rank therefore ends with bitwise the same 1024 floats, in one
 * launch, with one NVLink store latency + one flag latency on the critical path (4 KiB per peer:
 * the exchange is latency-, not bandwidth-bound, so a one-shot push beats a ring or tree).
 *
 * Mailbox (per rank, cudaMalloc'd, zeroed):   float slots[2][world][1024];  uint32 flags[2][world][4]
 * indexed by the parity of the call's sequence number: a rank can only reach call k+2 after it saw
 * every peer's flag of call k+1, i.e. after every peer finished reading the slots of call k.
 * The wait is bounded (clock64): a missing peer makes the call fail, it cannot hang the GPU; the output is
 * then NaN-filled.  After a failed call the ranks no longer agree on the sequence number / slot parity:
 * the exchange has to be destroyed and re-created on EVERY rank (b200sdr_exchange_destroy / _create / _connect).
 */
#ifndef B200_EXCHANGE_CUH
#define B200_EXCHANGE_CUH

#include <stdint.h>

#define B200_XCHG_MAX_WORLD 16
#define B200_XCHG_CTAS 4 /* 4 x 256 threads = 1024 bins */

#define B200_XCHG_SLOT_FLOATS(world) (2u * (world) * 1024u)
#define B200_XCHG_MAILBOX_BYTES(world) (B200_XCHG_SLOT_FLOATS(world) * 4u + 2u * (world) * B200_XCHG_CTAS * 4u)

struct ExchangeParams {
    const float *partials;      /* [ctas_per_capture][1024] of this rank's slice            */
    uint32_t ctas_per_capture;
    float scale;                /* 1 / frames of the WHOLE capture                          */
    float *out;                 /* 1024 floats, local                                       */
    float *mail[B200_XCHG_MAX_WORLD]; /* mailbox of every rank as mapped in THIS process    */
    uint32_t world, rank, seq;  /* seq >= 1, the same on every rank for the same call        */
    uint32_t *status;           /* device word, set to 1 when the wait timed out            */
    long long timeout_cycles;
};

#if defined(__CUDACC__) && !defined(B200_EMULATED)
__device__ __forceinline__ void b200_st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t b200_ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) k_spectrum_finalize_exchange(ExchangeParams p)
{
    __shared__ int s_timeout;
    const uint32_t cta = blockIdx.x, k = cta * 256u + threadIdx.x;
    const uint32_t par = p.seq & 1u;
    if (threadIdx.x == 0) s_timeout = 0;

    /* local reduction, fixed order (same arithmetic as k_spectrum_finalize) */
    float s = 0.0f;
    for (uint32_t i = 0; i < p.ctas_per_capture; ++i) s += p.partials[(uint64_t)i * 1024u + k];
    s *= p.scale;

    /* push into slot [par][rank] of every mailbox (own included): coalesced 1 KiB per CTA and peer */
    const uint32_t slot = (par * p.world + p.rank) * 1024u + k;
    for (uint32_t r = 0; r < p.world; ++r) p.mail[r][slot] = s;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < p.world) {
        uint32_t *flags = reinterpret_cast<uint32_t *>(p.mail[threadIdx.x] + B200_XCHG_SLOT_FLOATS(p.world));
        b200_st_release_sys(flags + (par * p.world + p.rank) * B200_XCHG_CTAS + cta, p.seq);
    }

    /* wait for this CTA's quarter from every rank */
    if (threadIdx.x < p.world) {
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(p.mail[p.rank] + B200_XCHG_SLOT_FLOATS(p.world)) +
                               (par * p.world + threadIdx.x) * B200_XCHG_CTAS + cta;
        const long long t0 = clock64();
        while (b200_ld_acquire_sys(mine) != p.seq) {
            if (clock64() - t0 > p.timeout_cycles) {
                s_timeout = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (s_timeout) { /* a peer never arrived: no stale spectrum is left behind, and the host sees the status word */
        p.out[k] = __int_as_float(0x7fc00000);
        if (threadIdx.x == 0) *p.status = 1u;
        return;
    }
    /* add the slots in rank order: identical bits on every rank.  The slots were written by peers over
     * NVLink: read them past L1 (it is not coherent with remote writes) */
    float sum = 0.0f;
    const float *slots = p.mail[p.rank] + par * p.world * 1024u + k;
    for (uint32_t r = 0; r < p.world; ++r) sum += __ldcg(slots + r * 1024u);
    p.out[k] = sum;
}
#endif

#endif
