/* TEST INFRASTRUCTURE: vdlm2dec_b200/csrc/vdl2_avlc.cuh (the per-frame walk of the row-f4 kernel) compiled for the host, so
   the SAME source can be compared byte for byte with the oracle's independent port before any GPU time is spent. */
#include <stddef.h>
#include <stdint.h>
#include "vdl2_avlc.cuh"

extern "C" void emul_avlc(const uint8_t * frames, int nframes, uint8_t * recs)
{				/* frames: vdl2_frame_t records (2048 B: len at offset 4, hdata at offset 32) */
	for (int f = 0; f < nframes; f++) {
		const uint8_t *fr = frames + (size_t) 2048 * f;
		int32_t l;
		__builtin_memcpy(&l, fr + 4, 4);
		if (l < 0)
			l = 0;
		if (l > 2016)
			l = 2016;
		avlc_extract(fr + 32, l, reinterpret_cast < Vdl2AvlcRec * >(recs + (size_t) 48 * f));
	}
}
