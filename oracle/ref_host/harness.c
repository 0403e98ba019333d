/*
 * harness.c -- oracle A: runs the UNMODIFIED sample-transfer path of the reference on the host.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Linked against the reference's own, unmodified translation units (compiled from
 * /root/reference by oracle/Makefile; nothing is copied into this repo):
 *   RTL/Src/usbh_rtlsdr.c            class: InterfaceInit (:163-260), USBH_RTLSDR_Process (:1058-1101)
 *   USBH/Src/usbh_ioreq.c            USBH_BulkReceiveData (:218-232)
 *   HAL_Driver/Src/stm32f7xx_hal_hcd.c   HAL_HCD_HC_SubmitRequest (:331-448), HAL_HCD_IRQHandler (:455-551),
 *                                        HCD_HC_IN_IRQHandler (:801-937), HCD_RXQLVL_IRQHandler (:1087-1132)
 *   RTL/Src/tuner_e4k.c              E4K_compute_pll_params (:689-737), for the front-end KATs
 *   HAL_Driver/Src/stm32f7xx_ll_usb.c    USB_HC_Init, USB_HC_StartXfer (:1426-1530), USB_ReadPacket (:792-803),
 *                                        USB_HC_Halt -- through ll_usb_shim.c (FIFO register redirected)
 *
 * What is NOT the reference and is written here: (1) the USBH<->HAL glue forwarders, restating
 * USBH/Src/usbh_conf.c:350-353, :423-441, :457-460 (the real file drags in GPIO/NVIC bring-up);
 * (2) stubs for the control plane the path never exercises (pipes, CtlReq, tuner, TIM HAL);
 * (3) the "hardware": two anonymous mappings at the MCU's physical addresses -- SDRAM
 * 0xC0000000 (8 MiB; the class puts its buffer at 0xC007F800, usbh_rtlsdr.c:227) and the
 * peripheral window 0x40000000 (1 MiB; TIM5 0x40000C00, RCC 0x40023800, OTG-HS 0x40040000) --
 * and a model of the OTG core that, for an armed IN channel, raises RXFLVL once per received
 * packet and XFRC / CHH at the end, by writing the registers the real IRQ handler reads.
 *
 * Exposed as a shared library (ref_read_packet needs no mappings and is safe inside python) and
 * as the CLI `ref_ingest_cli` (tests run the FSM in a subprocess so the fixed mappings cannot
 * collide with the python heap).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include "stm32f7xx_hal.h"
#include "usbh_core.h"
#include "usbh_rtlsdr.h"
#include "tuner_e4k.h"

/* ---- from ll_usb_shim.c ---- */
void ref_fifo_set_source(const uint8_t *src);
uint64_t ref_fifo_pops(void);
/* e4k_wrap.c: the reference's file-static E4000 selection routines */
int ref_e4k_rf_filter(int band, uint32_t freq);
int ref_e4k_if_bw_index(int filter, uint32_t bw);
uint32_t ref_e4k_if_bw_hz(int filter, int idx);
const uint8_t *ref_fifo_cursor(void);

/* ---------------------------------------------------------------------------------------------
 * 1. USB_ReadPacket alone (row a6).  Returns the number of bytes written to dest (whole words).
 * `src` must be readable up to the next multiple of 4 past len.
 * ------------------------------------------------------------------------------------------- */
uint32_t ref_read_packet(uint8_t *dest, const uint8_t *src, uint16_t len)
{
    uint64_t before = ref_fifo_pops();
    ref_fifo_set_source(src);
    uint8_t *end = (uint8_t *)USB_ReadPacket((USB_OTG_GlobalTypeDef *)0, dest, len);
    uint32_t written = (uint32_t)(end - dest);
    if ((ref_fifo_pops() - before) * 4u != written) return 0xFFFFFFFFu;
    return written;
}

/* A whole block through the reference's own copy routine, one max-size packet (512 bytes,
 * USB_EPA_MAXPKT, usbh_rtlsdr.c:830) at a time, as HCD_RXQLVL_IRQHandler feeds it (hal_hcd.c:1108-1112).
 * Thread-safe (the FIFO cursor is thread-local); used by bench.py's CPU baseline so that the part of
 * the path the reference DOES implement is timed with the reference's code.  Returns bytes written. */
size_t ref_copy_block(uint8_t *dest, const uint8_t *src, size_t nbytes)
{
    size_t off = 0;
    while (off < nbytes) {
        uint16_t n = (uint16_t)(nbytes - off < 512 ? nbytes - off : 512);
        ref_fifo_set_source(src + off);
        USB_ReadPacket((USB_OTG_GlobalTypeDef *)0, dest + off, n);
        off += n;
    }
    return (off + 3) & ~(size_t)3;
}

/* ---------------------------------------------------------------------------------------------
 * 2. stubs and glue
 * ------------------------------------------------------------------------------------------- */
uint32_t SystemCoreClock = 200000000u; /* src/main.c:240-258: 200 MHz */
/* Tuner_E4K comes from the reference's own tuner_e4k.c (linked for the PLL-parameter KATs) */

static HCD_HandleTypeDef g_hhcd;
static USBH_HandleTypeDef g_host;
static uint32_t g_submits;
static uint32_t g_tim_inits, g_tim_starts;

void HAL_Delay(__IO uint32_t Delay) { (void)Delay; }
HAL_StatusTypeDef HAL_TIM_Base_Init(TIM_HandleTypeDef *htim) { (void)htim; g_tim_inits++; return HAL_OK; }
HAL_StatusTypeDef HAL_TIM_Base_Start(TIM_HandleTypeDef *htim) { (void)htim; g_tim_starts++; return HAL_OK; }

uint8_t USBH_FindInterface(USBH_HandleTypeDef *phost, uint8_t Class, uint8_t SubClass, uint8_t Protocol)
{
    (void)phost; (void)Class; (void)SubClass; (void)Protocol;
    return 0;
}
USBH_StatusTypeDef USBH_SelectInterface(USBH_HandleTypeDef *phost, uint8_t interface)
{
    phost->device.current_interface = interface;
    return USBH_OK;
}
uint8_t USBH_AllocPipe(USBH_HandleTypeDef *phost, uint8_t ep_addr) { (void)phost; (void)ep_addr; return 2; }
USBH_StatusTypeDef USBH_FreePipe(USBH_HandleTypeDef *phost, uint8_t idx) { (void)phost; (void)idx; return USBH_OK; }
USBH_StatusTypeDef USBH_ClosePipe(USBH_HandleTypeDef *phost, uint8_t pipe_num) { (void)phost; (void)pipe_num; return USBH_OK; }
/* restates USBH_OpenPipe -> USBH_LL_OpenPipe -> HAL_HCD_HC_Init (usbh_pipes.c:93-112, usbh_conf.c:384-398) */
USBH_StatusTypeDef USBH_OpenPipe(USBH_HandleTypeDef *phost, uint8_t pipe_num, uint8_t epnum, uint8_t dev_address,
                                 uint8_t speed, uint8_t ep_type, uint16_t mps)
{
    HAL_HCD_HC_Init((HCD_HandleTypeDef *)phost->pData, pipe_num, epnum, dev_address, speed, ep_type, mps);
    return USBH_OK;
}
/* Control transfers complete at once.  While a trace is being recorded (section 5) every transfer is
 * logged as the setup packet the reference put into phost->Control.setup plus its payload; IN transfers are
 * answered by a two-line device model: an I2C read through the repeater returns the E4000's chip id when
 * register E4K_CHECK_ADDR of E4K_I2C_ADDR was addressed last (so the reference's own probe finds a tuner,
 * usbh_rtlsdr.c:961-975), everything else reads 0. */
struct ref_ctl_rec { uint8_t bmRequestType, bRequest; uint16_t wValue, wIndex, wLength; uint8_t data[2]; uint8_t step, pad; };
static struct ref_ctl_rec *g_trace;
static uint32_t g_trace_cap, g_trace_n;
static uint16_t g_i2c_last_addr;
static uint8_t g_i2c_last_reg;
static uint8_t g_e4k_regs[256]; /* E4000 register file: a write of (reg, val) stores, a write of (reg) selects, a read returns */
static int g_full_tuner;        /* 1: keep the reference's own E4000 driver in the loop (ref_init_trace_full) */
static uint8_t ref_trace_step(void);
USBH_StatusTypeDef USBH_CtlReq(USBH_HandleTypeDef *phost, uint8_t *buff, uint16_t length)
{
    if (!g_trace) return USBH_OK;
    const USB_Setup_TypeDef *su = &phost->Control.setup;
    const int in = (su->b.bmRequestType & 0x80) != 0;
    if (in) {
        memset(buff, 0, length);
        if (su->b.wIndex.w == (IICB << 8) && su->b.wValue.w == E4K_I2C_ADDR && g_i2c_last_addr == E4K_I2C_ADDR && length >= 1)
            buff[0] = g_i2c_last_reg == E4K_CHECK_ADDR ? E4K_CHECK_VAL : g_e4k_regs[g_i2c_last_reg];
    } else if (su->b.wIndex.w == ((IICB << 8) | 0x10) && length >= 1) {
        g_i2c_last_addr = su->b.wValue.w;
        g_i2c_last_reg = buff[0];
        if (su->b.wValue.w == E4K_I2C_ADDR && length >= 2) g_e4k_regs[buff[0]] = buff[1];
    }
    if (g_trace_n < g_trace_cap) {
        struct ref_ctl_rec *r = &g_trace[g_trace_n];
        memset(r, 0, sizeof *r);
        r->bmRequestType = su->b.bmRequestType;
        r->bRequest = su->b.bRequest;
        r->wValue = su->b.wValue.w;
        r->wIndex = su->b.wIndex.w;
        r->wLength = su->b.wLength.w;
        if (!in) memcpy(r->data, buff, length < 2 ? length : 2);
        r->step = ref_trace_step();
    }
    g_trace_n++;
    return USBH_OK;
}
USBH_StatusTypeDef USBH_Process(USBH_HandleTypeDef *phost) { (void)phost; return USBH_OK; }

/* glue: usbh_conf.c:423-441 */
USBH_StatusTypeDef USBH_LL_SubmitURB(USBH_HandleTypeDef *phost, uint8_t pipe, uint8_t direction, uint8_t ep_type,
                                     uint8_t token, uint8_t *pbuff, uint16_t length, uint8_t do_ping)
{
    g_submits++;
    HAL_HCD_HC_SubmitRequest((HCD_HandleTypeDef *)phost->pData, pipe, direction, ep_type, token, pbuff, length, do_ping);
    return USBH_OK;
}
/* glue: usbh_conf.c:457-460 */
USBH_URBStateTypeDef USBH_LL_GetURBState(USBH_HandleTypeDef *phost, uint8_t pipe)
{
    return (USBH_URBStateTypeDef)HAL_HCD_HC_GetURBState((HCD_HandleTypeDef *)phost->pData, pipe);
}
/* glue: usbh_conf.c:350-353 */
uint32_t USBH_LL_GetLastXferSize(USBH_HandleTypeDef *phost, uint8_t pipe)
{
    return HAL_HCD_HC_GetXferCount((HCD_HandleTypeDef *)phost->pData, pipe);
}
/* glue: usbh_conf.c:479-491 */
USBH_StatusTypeDef USBH_LL_SetToggle(USBH_HandleTypeDef *phost, uint8_t pipe, uint8_t toggle)
{
    HCD_HandleTypeDef *h = (HCD_HandleTypeDef *)phost->pData;
    if (h->hc[pipe].ep_is_in) h->hc[pipe].toggle_in = toggle;
    else h->hc[pipe].toggle_out = toggle;
    return USBH_OK;
}

/* ---------------------------------------------------------------------------------------------
 * 3. the "board": memory at the MCU's physical addresses
 * ------------------------------------------------------------------------------------------- */
#ifndef MAP_FIXED_NOREPLACE
#define MAP_FIXED_NOREPLACE 0x100000
#endif
static int g_env_ready;

int ref_env_init(void)
{
    if (g_env_ready) return 0;
    void *a = mmap((void *)0xC0000000ul, 8u << 20, PROT_READ | PROT_WRITE,
                   MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED_NOREPLACE, -1, 0);
    if (a != (void *)0xC0000000ul) return -1;
    void *b = mmap((void *)0x40000000ul, 1u << 20, PROT_READ | PROT_WRITE,
                   MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED_NOREPLACE, -1, 0);
    if (b != (void *)0x40000000ul) return -2;
    g_env_ready = 1;
    return 0;
}

/* Class Init through the registered vtable, as USBH_Process does at usbh_core.c:527. */
int ref_class_init(void)
{
    if (ref_env_init() != 0) return -1;
    memset(&g_hhcd, 0, sizeof g_hhcd);
    memset(&g_host, 0, sizeof g_host);
    /* usbh_conf.c:244-256 */
    g_hhcd.Instance = USB_OTG_HS;
    g_hhcd.Init.Host_channels = 11;
    g_hhcd.Init.dma_enable = 0;
    g_hhcd.Init.speed = HCD_SPEED_HIGH;
    g_hhcd.pData = &g_host;
    g_host.pData = &g_hhcd;
    USB_OTG_HS->GINTSTS = 1u;        /* CMOD = host, so USB_GetMode() says host */
    USB_OTG_HS->HNPTXSTS = 0x00080100u; /* request queue / FIFO space available: USB_HC_Halt takes the short path */
    /* what enumeration would have left: EP 0x81 bulk, 512-byte packets (USB_EPA_MAXPKT, usbh_rtlsdr.c:830) */
    g_host.device.address = 1;
    g_host.device.speed = USBH_SPEED_HIGH;
    g_host.device.is_connected = 1;
    g_host.device.CfgDesc.Itf_Desc[0].Ep_Desc[0].bEndpointAddress = 0x81;
    g_host.device.CfgDesc.Itf_Desc[0].Ep_Desc[0].wMaxPacketSize = 512;
    g_host.pClass[0] = USBH_RTLSDR_CLASS;
    g_host.ClassNumber = 1;
    g_host.pActiveClass = USBH_RTLSDR_CLASS;
    g_host.gState = HOST_CLASS;
    g_submits = 0;
    return (int)g_host.pActiveClass->Init(&g_host);
}

static RTLSDR_HandleTypeDef *handle(void) { return (RTLSDR_HandleTypeDef *)g_host.pActiveClass->pData; }

uint64_t ref_class_buff_addr(void) { return (uint64_t)(uintptr_t)handle()->CommItf.buff; }
uint32_t ref_class_buff_size(void) { return handle()->CommItf.buffSize; }
uint32_t ref_class_ep(void) { return handle()->CommItf.SdrEp; }
uint32_t ref_class_mps(void) { return handle()->CommItf.SdrEpSize; }
uint32_t ref_class_xfer_state(void) { return (uint32_t)handle()->xferState; }
uint32_t ref_class_wait_polls(void) { return handle()->xferWaitNo; }
uint32_t ref_submit_count(void) { return g_submits; }
uint32_t ref_tim_prescaler(void) { return handle()->uwPrescalerValue; }
/* "This should be user configurable" (usbh_rtlsdr.c:229): the one knob the harness turns */
void ref_class_set_buff_size(uint32_t n) { handle()->CommItf.buffSize = n; }
void ref_tim5_set_cnt(uint32_t v) { TIM5->CNT = v; }
uint32_t ref_tim5_get_cnt(void) { return TIM5->CNT; }
int ref_class_bgnd(void) { return (int)g_host.pActiveClass->BgndProcess(&g_host); }
int ref_class_deinit(void) { return (int)g_host.pActiveClass->DeInit(&g_host); }
const uint8_t *ref_class_buff(void) { return handle()->CommItf.buff; }

/* ---------------------------------------------------------------------------------------------
 * OTG core model.  For the armed IN channel: deliver `avail` stream bytes as max-packet-size
 * packets through the real IRQ handler until HCTSIZ.PKTCNT reaches 0 (or a short packet ends the
 * transfer), then XFRC and CHH.  Returns bytes consumed from the stream, or <0 on model error.
 * ------------------------------------------------------------------------------------------- */
static void raise_irq(uint32_t gintsts_flag)
{
    USB_OTG_HS->GINTMSK |= gintsts_flag;
    USB_OTG_HS->GINTSTS = 1u | gintsts_flag;
    HAL_HCD_IRQHandler(&g_hhcd);
    USB_OTG_HS->GINTSTS = 1u;
}

long ref_hw_deliver(const uint8_t *stream, size_t avail)
{
    uint8_t ch = handle()->CommItf.SdrPipe;
    USB_OTG_GlobalTypeDef *USBx = USB_OTG_HS;
    if (!(USBx_HC(ch)->HCCHAR & USB_OTG_HCCHAR_CHENA)) return -1; /* not armed */
    uint32_t mps = g_hhcd.hc[ch].max_packet;
    size_t used = 0;
    for (;;) {
        uint32_t hctsiz = USBx_HC(ch)->HCTSIZ;
        uint32_t pktcnt = (hctsiz & USB_OTG_HCTSIZ_PKTCNT) >> 19;
        uint32_t xfrsiz = hctsiz & USB_OTG_HCTSIZ_XFRSIZ;
        if (pktcnt == 0) break;
        uint32_t n = mps;
        if (n > avail - used) n = (uint32_t)(avail - used);
        /* the core decrements the packet count and remaining size as it receives */
        pktcnt -= 1;
        xfrsiz = xfrsiz >= n ? xfrsiz - n : 0;
        USBx_HC(ch)->HCTSIZ = (hctsiz & USB_OTG_HCTSIZ_DPID) | (pktcnt << 19) | xfrsiz;
        /* receive-status word popped by the handler: channel, byte count, PKTSTS = IN data */
        USBx->GRXSTSP = (uint32_t)ch | (n << 4) | ((uint32_t)GRXSTS_PKTSTS_IN << 17);
        ref_fifo_set_source(stream + used);
        raise_irq(USB_OTG_GINTSTS_RXFLVL);
        used += n;
        if (n < mps) break; /* short packet terminates the transfer */
    }
    /* transfer complete on the channel, then channel halted */
    USBx_HOST->HAINT = 1u << ch;
    USBx_HC(ch)->HCINT = USB_OTG_HCINT_XFRC;
    raise_irq(USB_OTG_GINTSTS_HCINT);
    USBx_HC(ch)->HCCHAR &= ~USB_OTG_HCCHAR_CHENA; /* halt request honoured by the core */
    USBx_HC(ch)->HCINT = USB_OTG_HCINT_CHH;
    raise_irq(USB_OTG_GINTSTS_HCINT);
    USBx_HOST->HAINT = 0;
    return (long)used;
}

/* Run the whole stream through the class FSM exactly as the superloop would
 * (src/main.c:72-80 -> usbh_core.c:575-581): BgndProcess is called again and again; when it is
 * waiting, the core model delivers the next URB's worth of the stream; when it reports
 * RTLSDR_XFER_COMPLETE the buffer content is appended to `out`.  Returns blocks completed. */
long ref_run_stream(const uint8_t *stream, size_t total, uint32_t buff_size, uint32_t tim_cnt, uint8_t *out,
                    uint32_t *polls_per_block, uint32_t max_blocks)
{
    if (ref_class_init() != 0) return -1;
    ref_class_set_buff_size(buff_size);
    size_t off = 0;
    uint32_t blocks = 0;
    while (off < total && blocks < max_blocks) {
        uint32_t calls = 0;
        /* START: submits the URB */
        if (ref_class_xfer_state() != RTLSDR_XFER_START) return -2;
        ref_class_bgnd(); calls++;
        if (ref_class_xfer_state() != RTLSDR_XFER_WAIT) return -3;
        long used = ref_hw_deliver(stream + off, total - off);
        if (used < 0) return -4;
        ref_tim5_set_cnt(tim_cnt);
        ref_class_bgnd(); calls++; /* WAIT: sees URB_DONE, logs, -> COMPLETE */
        if (ref_class_xfer_state() != RTLSDR_XFER_COMPLETE) return -5;
        /* <- this is where process_samples(buff, buffSize, ctx) hooks in */
        memcpy(out + off, ref_class_buff(), (size_t)used);
        ref_class_bgnd(); calls++; /* COMPLETE -> START */
        if (polls_per_block) polls_per_block[blocks] = calls;
        off += (size_t)used;
        blocks++;
    }
    return (long)blocks;
}

/* ---------------------------------------------------------------------------------------------
 * 4. front-end parameter math of the reference (section 8f rows 2 and 4), called as the firmware
 * calls it: the first state of each routine does the arithmetic and stores it in the handle.
 * ------------------------------------------------------------------------------------------- */
int ref_set_sample_rate(uint32_t rate, uint32_t *rsamp_ratio, uint32_t *real_rsamp_ratio, double *real_rate)
{
    if (ref_class_init() != 0) return -1;
    handle()->setSampleRateState = 0;
    RTLSDR_set_sample_rate(&g_host, rate); /* usbh_rtlsdr.c:666-700, state 0 */
    *rsamp_ratio = handle()->rsamp_ratio;
    *real_rsamp_ratio = handle()->real_rsamp_ratio;
    *real_rate = handle()->real_rate;
    return 0;
}

int ref_fir_bytes(uint8_t out20[20])
{
    if (ref_class_init() != 0) return -1;
    handle()->firState = RTLSDR_FIR_CALC;
    RTLSDR_set_fir(&g_host); /* usbh_rtlsdr.c:534-575, state RTLSDR_FIR_CALC */
    memcpy(out20, handle()->fir, 20);
    return 0;
}

/* out8: fosc, intended_flo, flo, x, z, r, r_idx, threephase */
uint32_t ref_e4k_pll(uint32_t fosc, uint32_t intended_flo, uint32_t out8[8])
{
    struct e4k_pll_params p;
    memset(&p, 0, sizeof p);
    uint32_t flo = E4K_compute_pll_params(&p, fosc, intended_flo); /* tuner_e4k.c:689-737 */
    out8[0] = p.fosc; out8[1] = p.intended_flo; out8[2] = p.flo; out8[3] = p.x;
    out8[4] = p.z; out8[5] = p.r; out8[6] = p.r_idx; out8[7] = p.threephase;
    return flo;
}

/* ---------------------------------------------------------------------------------------------
 * 5. the register-initialisation sequence (section 8f row 2, "register sequences as data"): the reference's
 * own, unmodified USBH_RTLSDR_ClassRequest (usbh_rtlsdr.c:809-945), reached through the class vtable as
 * USBH_Process does (usbh_core.c:557), polled until it reports the class active.  Every control transfer it
 * issues is recorded by USBH_CtlReq above.  The E4000 driver's own init (steps 28, 29 and the SetBW call
 * inside RTLSDR_set_sample_rate) is a separate state machine over dozens of I2C registers; it is replaced by
 * a tuner that accepts everything, after the reference's probe has found "its" E4000 (ref_init_trace_full keeps
 * the reference's driver instead, against a register-file model of the chip).
 * Returns the number of transfers (which may exceed cap), or < 0.
 * ------------------------------------------------------------------------------------------- */
static USBH_StatusTypeDef tuner_ok(USBH_HandleTypeDef *phost) { (void)phost; return USBH_OK; }
static RTLSDR_TunerTypeDef g_null_tuner;
static void user_cb(USBH_HandleTypeDef *phost, uint8_t id) { (void)phost; (void)id; }
static uint8_t ref_trace_step(void) { return handle()->reqNumber; }

long ref_init_trace(uint8_t *out, uint32_t cap_records)
{
    if (ref_class_init() != 0) return -1;
    g_null_tuner.Name = "none";
    g_null_tuner.Init = tuner_ok;
    g_null_tuner.InitProcess = tuner_ok;
    g_null_tuner.SetBW = tuner_ok;
    g_host.RequestState = CMD_SEND; /* what USBH_CtlReq leaves between requests (usbh_ctlreq.c) */
    g_host.pUser = user_cb;
    g_trace = (struct ref_ctl_rec *)out;
    g_trace_cap = cap_records;
    g_trace_n = 0;
    g_i2c_last_addr = 0;
    g_i2c_last_reg = 0;
    memset(g_e4k_regs, 0, sizeof g_e4k_regs);
    long rc = -2;
    for (int polls = 0; polls < 400000; ++polls) {
        if (!g_full_tuner && handle()->reqNumber == 28 && handle()->tuner == &Tuner_E4K) handle()->tuner = &g_null_tuner;
        USBH_StatusTypeDef st = g_host.pActiveClass->Requests(&g_host);
        if (st == USBH_OK) { rc = (long)g_trace_n; break; }
        /* anything else means "call me again": USBH_Process only looks for USBH_OK (usbh_core.c:557-562), and the
         * FSM's own return value is USBH_FAIL while a sub-FSM is busy (rStatus is never set in that branch) */
    }
    g_trace = 0;
    return rc;
}

/* the same with the reference's own E4000 driver left in (steps 28, 29, SetBW): does its init FSM terminate against
 * a plain register-file model of the chip, and what does it put on the wire? */
long ref_init_trace_full(uint8_t *out, uint32_t cap_records)
{
    g_full_tuner = 1;
    long n = ref_init_trace(out, cap_records);
    g_full_tuner = 0;
    return n;
}

#ifdef REF_CLI
static int cli_frontend(int argc, char **argv)
{
    if (strcmp(argv[1], "--init-trace-full") == 0 && argc >= 3) {
        static uint8_t buf[12 * 8192];
        long n = ref_init_trace_full(buf, 8192);
        printf("REF_INIT_TRACE_FULL %ld step=%u\n", n, (unsigned)handle()->reqNumber);
        if (n < 0 || n > 8192) return 1;
        FILE *f = fopen(argv[2], "wb");
        if (!f) return 1;
        fwrite(buf, 12, (size_t)n, f);
        fclose(f);
        return 0;
    }
    if (strcmp(argv[1], "--init-trace") == 0 && argc >= 3) { /* writes the 12-byte records to argv[2] */
        static uint8_t buf[12 * 4096];
        long n = ref_init_trace(buf, 4096);
        if (n < 0 || n > 4096) { fprintf(stderr, "init trace failed: %ld\n", n); return 1; }
        FILE *f = fopen(argv[2], "wb");
        if (!f) return 1;
        fwrite(buf, 12, (size_t)n, f);
        fclose(f);
        printf("REF_INIT_TRACE %ld\n", n);
        return 0;
    }
    if (strcmp(argv[1], "--rate") == 0 && argc >= 3) {
        uint32_t a, b; double r;
        if (ref_set_sample_rate((uint32_t)strtoul(argv[2], 0, 0), &a, &b, &r)) return 1;
        printf("REF_RATE %u %u %.17g\n", a, b, r);
        return 0;
    }
    if (strcmp(argv[1], "--fir") == 0) {
        uint8_t f[20];
        if (ref_fir_bytes(f)) return 1;
        printf("REF_FIR ");
        for (int i = 0; i < 20; ++i) printf("%02x", f[i]);
        printf("\n");
        return 0;
    }
    if (strcmp(argv[1], "--e4k") == 0 && argc >= 4) {
        uint32_t o[8];
        uint32_t flo = ref_e4k_pll((uint32_t)strtoul(argv[2], 0, 0), (uint32_t)strtoul(argv[3], 0, 0), o);
        printf("REF_E4K %u %u %u %u %u %u %u %u %u\n", flo, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
        return 0;
    }
    if (strcmp(argv[1], "--e4k-rf") == 0 && argc >= 4) { /* e4k_wrap.c -> choose_rf_filter */
        printf("REF_E4K_RF %d\n", ref_e4k_rf_filter(atoi(argv[2]), (uint32_t)strtoul(argv[3], 0, 0)));
        return 0;
    }
    if (strcmp(argv[1], "--e4k-ifbw") == 0 && argc >= 4) { /* e4k_wrap.c -> find_if_bw */
        int idx = ref_e4k_if_bw_index(atoi(argv[2]), (uint32_t)strtoul(argv[3], 0, 0));
        printf("REF_E4K_IFBW %d %u\n", idx, ref_e4k_if_bw_hz(atoi(argv[2]), idx));
        return 0;
    }
    return -1;
}

/* ref_ingest_cli <in.bin> <buff_size> <tim_cnt> <out.bin>
 * stdout: the reference's own log lines, then one summary line starting with "REF_SUMMARY". */
int main(int argc, char **argv)
{
    if (argc >= 2 && argv[1][0] == '-') {
        int rc = cli_frontend(argc, argv);
        if (rc >= 0) return rc;
    }
    if (argc < 5) { fprintf(stderr, "usage: %s in.bin buff_size tim_cnt out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END);
    long total = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *in = (uint8_t *)calloc((size_t)total + 1024, 1);
    uint8_t *out = (uint8_t *)calloc((size_t)total + 1024, 1);
    if (fread(in, 1, (size_t)total, f) != (size_t)total) return 4;
    fclose(f);
    uint32_t buff_size = (uint32_t)strtoul(argv[2], 0, 0), tim_cnt = (uint32_t)strtoul(argv[3], 0, 0);
    uint32_t *polls = (uint32_t *)calloc(1u << 20, sizeof(uint32_t));
    int st = ref_class_init();
    printf("REF_INIT status=%d buff=0x%llx size=%u ep=0x%02x mps=%u prescaler=%u\n", st,
           (unsigned long long)ref_class_buff_addr(), ref_class_buff_size(), ref_class_ep(), ref_class_mps(),
           ref_tim_prescaler());
    long blocks = ref_run_stream(in, (size_t)total, buff_size, tim_cnt, out, polls, 1u << 20);
    uint32_t pmin = 0xFFFFFFFFu, pmax = 0;
    for (long b = 0; b < blocks; ++b) { if (polls[b] < pmin) pmin = polls[b]; if (polls[b] > pmax) pmax = polls[b]; }
    printf("REF_SUMMARY blocks=%ld submits=%u polls_min=%u polls_max=%u buff=0x%llx size=%u final_state=%u\n", blocks,
           ref_submit_count(), pmin, pmax, (unsigned long long)ref_class_buff_addr(), ref_class_buff_size(),
           ref_class_xfer_state());
    f = fopen(argv[4], "wb");
    if (!f) return 5;
    fwrite(out, 1, (size_t)total, f);
    fclose(f);
    return blocks < 0 ? 1 : 0;
}
#endif
