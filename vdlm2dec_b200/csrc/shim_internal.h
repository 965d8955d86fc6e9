/* shim_internal.h -- the two entry points d8psk_shim.c offers to the other front ends of the drop-in
   (its own rcv_thread loop behind rtl.c / air.c, and file_shim.c for row f2).  Not part of the C ABI. */
#ifndef VDL2_SHIM_INTERNAL_H
#define VDL2_SHIM_INTERNAL_H
#include <stddef.h>

/* number of channels main.c registered (main.c:59) */
int vdl2shim_nch(void);
/* create the GPU handle for the registered channels, all demodulated from ONE input stream;
   max_samples = largest nsamples of a vdl2shim_feed() call.  exit(1) on failure (main.c:209-213). */
void vdl2shim_open(unsigned fs, unsigned sdrclk, int format, size_t max_samples);
/* demodulate nsamples samples of the stream (host memory, format of vdl2shim_open) and deliver what completed:
   blocks to decodeVdlm2() (vdlm2.c:189-205), or with -DVDL2_SHIM_LINK frames to out() (vdlm2.h:134) */
void vdl2shim_feed(const void *iq, size_t nsamples);
/* same for raw cu8 bytes of whole 65536-byte RTL callbacks, demodulated as the reference's in_callback lays them out
   (rtl.c:285-292; expanded on the device, vdl2_process_host_rtl); the handle must have been opened with VDL2_FMT_CF32 */
void vdl2shim_feed_rtl(const void *cu8, size_t nsamples);
/* end of input: returns once what vdl2shim_feed() handed over has had time to reach out() */
void vdl2shim_finish(void);
#endif
