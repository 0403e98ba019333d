/*
 * ll_usb_shim.c -- compiles the UNMODIFIED reference HAL_Driver/Src/stm32f7xx_ll_usb.c for the
 * host, with the one hardware access that cannot exist here -- the OTG data FIFO register --
 * redirected to a byte stream supplied by the harness.
 *
 * TEST INFRASTRUCTURE ONLY (oracle A).  No reference source is copied: the file is #included
 * from where it lies (REF_LL_USB_C is passed by oracle/Makefile as an absolute path).
 *
 * USBx_DFIFO is defined at HAL_Driver/Inc/stm32f7xx_ll_usb.h:381 as a fixed register address;
 * on hardware every read pops one 32-bit word of the received packet.  USB_ReadPacket
 * (stm32f7xx_ll_usb.c:792-803) reads it (len+3)/4 times.
 */
#include <stdint.h>
#include <string.h>

#include "stm32f7xx_hal.h"

/* thread-local: bench.py's CPU baseline runs the reference copy on every host core at once */
static __thread const uint8_t *g_fifo_src; /* next byte the "FIFO" will deliver */
static __thread uint64_t g_fifo_pops;

void ref_fifo_set_source(const uint8_t *src) { g_fifo_src = src; }
uint64_t ref_fifo_pops(void) { return g_fifo_pops; }
const uint8_t *ref_fifo_cursor(void) { return g_fifo_src; }

/* one FIFO pop = the next whole 32-bit word of the stream (the tail of a short packet is padding).
 * Kept inline so the copy loop of USB_ReadPacket costs what a register read would, not a call. */
typedef uint32_t __attribute__((aligned(1), may_alias)) ref_u32_unaligned;
static inline uint32_t ref_fifo_pop(void)
{
    uint32_t w = *(const ref_u32_unaligned *)g_fifo_src;
    g_fifo_src += 4;
    g_fifo_pops++;
    return w;
}
static __thread uint32_t g_fifo_sink; /* target of FIFO writes (USB_WritePacket is never exercised) */

#undef USBx_DFIFO
#define USBx_DFIFO(i) (*((void)(i), (g_fifo_sink = ref_fifo_pop()), (volatile uint32_t *)&g_fifo_sink))

#include REF_LL_USB_C
