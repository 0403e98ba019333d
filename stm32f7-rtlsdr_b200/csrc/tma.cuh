/*
 * tma.cuh -- 1-D bulk asynchronous copies global -> shared (cp.async.bulk, SASS UBLKCP) with
 * mbarrier completion: the TMA engine moves the raw u8 tiles, no thread touches the bytes on
 * their way into shared memory and no registers are tied up.
 * Host emulation (tests/emu): synchronous memcpy, barriers are no-ops.
 */
#ifndef B200_TMA_CUH
#define B200_TMA_CUH

#include "cplx2.cuh"

#ifdef B200_PACKED /* real device code */

B200_DEV uint32_t b200_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

B200_DEV void b200_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b200_smem_u32(bar)), "r"(count));
}
B200_DEV void b200_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
B200_DEV void b200_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b200_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
B200_DEV void b200_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b200_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
/* bytes: multiple of 16; src and dst 16-byte aligned */
B200_DEV void b200_tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     b200_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(b200_smem_u32(bar))
                 : "memory");
}

#else /* host emulation */

#include <string.h>
B200_DEV void b200_mbar_init(uint64_t *, uint32_t) {}
B200_DEV void b200_mbar_fence_init() {}
B200_DEV void b200_mbar_expect_tx(uint64_t *, uint32_t) {}
B200_DEV void b200_mbar_wait(uint64_t *, uint32_t) {}
B200_DEV void b200_tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *)
{
    memcpy(smem_dst, gsrc, bytes);
}

#endif
#endif
