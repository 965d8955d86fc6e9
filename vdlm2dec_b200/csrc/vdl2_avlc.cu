/* vdl2_avlc.cu -- row f4: frames -> field records on the device.  One warp per frame: the 32 lanes copy the frame into
   shared memory with coalesced 16-byte loads (a frame record is 2048 bytes, 16-byte aligned), lane 0 then walks it
   (vdl2_avlc.cuh) and stores the 48-byte record.  Algorithmic traffic 2048 B in + 48 B out per frame; the walk is a
   dependent chain of at most ~2000 CRC steps, so the kernel is latency bound and sized by the number of frames (a few
   thousand per front-end step), not by HBM. */
#include <cuda_runtime.h>
#include "vdl2_avlc.cuh"
#include "vdl2_link.h"

#define AVLC_WARPS 4

__global__ void __launch_bounds__(32 * AVLC_WARPS) vdl2_avlc_kernel(const Vdl2FrameRec * __restrict__ frames, int nframes, Vdl2AvlcRec * __restrict__ recs)
{
	__shared__ uint4 stage[AVLC_WARPS][sizeof(Vdl2FrameRec) / 16];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int f = blockIdx.x * AVLC_WARPS + warp;
	if (f >= nframes)
		return;
	const uint4 *src = reinterpret_cast < const uint4 * >(frames + f);
	for (int i = lane; i < (int)(sizeof(Vdl2FrameRec) / 16); i += 32)
		stage[warp][i] = src[i];
	__syncwarp();
	if (lane == 0) {
		const Vdl2FrameRec *fr = reinterpret_cast < const Vdl2FrameRec * >(&stage[warp][0]);
		int l = fr->len;
		if (l < 0)
			l = 0;
		if (l > (int)sizeof fr->hdata)
			l = (int)sizeof fr->hdata;
		avlc_extract(fr->hdata, l, recs + f);
	}
}

extern "C" int vdl2_avlc_launch(const Vdl2FrameRec * d_frames, int nframes, void *d_recs, void *stream)
{
	if (nframes <= 0)
		return 0;
	const int grid = (nframes + AVLC_WARPS - 1) / AVLC_WARPS;
	vdl2_avlc_kernel <<< grid, 32 * AVLC_WARPS, 0, (cudaStream_t) stream >>> (d_frames, nframes, (Vdl2AvlcRec *) d_recs);
	return (int)cudaGetLastError();
}
