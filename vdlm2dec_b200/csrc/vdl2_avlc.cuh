/*
 * vdl2_avlc.cuh -- per-frame field extraction: what the reference's out() and outacars()
 * derive from the bytes of a frame before they format text or JSON.
 *   out.c:519-535   command/response bit, addresses (icaoaddr, out.c:426-435), direction, on-ground bit
 *   out.c:562-570   payload class: XID group (0x82), ACARS (ff ff 01), other, none
 *   outacars.c:214-290  ACARS: CRC over the body, parity strip, mode, registration, ack, label, block id, message number,
 *                   flight id, text extent, end-of-block character
 * Formatting (fixreg, label decoding, text, JSON) stays on the host.  The same source is compiled for the device
 * (vdl2_avlc.cu: a warp stages the frame in shared memory, computes the ACARS CRC lane-parallel -- it is linear -- and lane 0
 * walks the handful of header fields) and for the host test build
 * (tests/emul), where it is compared byte for byte with the oracle's independent port.
 */
#ifndef VDL2_AVLC_CUH
#define VDL2_AVLC_CUH
#include <stdint.h>

#ifdef __CUDACC__
#define VDL2_AVLC_FN __host__ __device__ __forceinline__
#else
#define VDL2_AVLC_FN static inline
#endif

/* identical layout to vdl2_avlc_t (include/vdl2gpu.h) */
struct Vdl2AvlcRec {
	uint32_t faddr, taddr;
	uint8_t fromair, rep, gnd, lc, kind;
	uint8_t mode, ack, bid, bs, be;
	uint8_t label[2];
	uint8_t reg[7];
	uint8_t nno, nfid;
	uint8_t no[4], fid[6];
	uint8_t pad;
	uint16_t txt_off, txt_len, info_off, info_len;
};
static_assert(sizeof(Vdl2AvlcRec) == 48, "vdl2_avlc_t layout");

enum { AVLC_EMPTY = 0, AVLC_XID = 1, AVLC_ACARS = 2, AVLC_ACARS_BADCRC = 3, AVLC_OTHER = 4 };

VDL2_AVLC_FN uint32_t avlc_rev8(uint32_t b)
{
#ifdef __CUDA_ARCH__
	return __brev(b) >> 24;
#else
	b = ((b & 0xf0u) >> 4) | ((b & 0x0fu) << 4);
	b = ((b & 0xccu) >> 2) | ((b & 0x33u) << 2);
	return ((b & 0xaau) >> 1) | ((b & 0x55u) << 1);
#endif
}

/* the four address octets carry 6 + 7 + 7 + 7 bits above their low marker bits, least significant first: reversing an octet
   puts its group in the low bits, most significant first (out.c:426-435 does the same with reversebits per group) */
VDL2_AVLC_FN uint32_t avlc_addr(const uint8_t * a)
{
	return ((avlc_rev8(a[0]) & 0x3fu) << 21) | ((avlc_rev8(a[1]) & 0x7fu) << 14) | ((avlc_rev8(a[2]) & 0x7fu) << 7) | (avlc_rev8(a[3]) & 0x7fu);
}

/* one octet into the reflected CRC-16/CCITT (crc.h:3 with crc.c's table): closed form of the table entry */
VDL2_AVLC_FN uint32_t avlc_crc(uint32_t crc, uint32_t c)
{
	uint32_t d = (c ^ crc) & 0xffu;
	d = (d ^ (d << 4)) & 0xffu;
	return ((crc >> 8) ^ (d << 8) ^ (d << 3) ^ (d >> 4)) & 0xffffu;
}

/* does the frame carry an ACARS body whose CRC has to be computed?  (out.c:566: ff ff 01 in front of it) */
VDL2_AVLC_FN bool avlc_is_acars(const uint8_t * h, int l)
{
	return l >= 16 && h[10] != 0x82 && h[10] == 0xff && h[11] == 0xff && h[12] == 0x01;
}

/* the field walk proper.  `crc` = CRC over the ACARS body t[0 .. n-2] (outacars.c:222-230), used only when avlc_is_acars();
   the device computes it lane-parallel (vdl2_avlc.cu), the host build and single-lane callers through avlc_extract() below */
VDL2_AVLC_FN void avlc_fields(const uint8_t * h, int l, uint32_t crc, Vdl2AvlcRec * r)
{
	Vdl2AvlcRec o;
	o.faddr = avlc_addr(h + 5);
	o.taddr = avlc_addr(h + 1);
	o.fromair = (o.faddr >> 24) == 1u;
	o.rep = (h[5] >> 1) & 1;
	o.gnd = (h[1] >> 1) & 1;
	o.lc = h[9];
	o.kind = AVLC_EMPTY;
	o.mode = o.ack = o.bid = o.bs = o.be = 0;
	o.label[0] = o.label[1] = 0;
	for (int i = 0; i < 7; i++)
		o.reg[i] = 0;
	o.nno = o.nfid = 0;
	for (int i = 0; i < 4; i++)
		o.no[i] = 0;
	for (int i = 0; i < 6; i++)
		o.fid[i] = 0;
	o.pad = 0;
	o.txt_off = o.txt_len = o.info_off = o.info_len = 0;
	if (l > 13) {
		o.info_off = 10;
		o.info_len = (uint16_t) (l - 13);
		if (h[10] == 0x82)
			o.kind = AVLC_XID;
		else if (l >= 16 && h[10] == 0xff && h[11] == 0xff && h[12] == 0x01) {
			const uint8_t *t = h + 13;
			const int n = l - 16;	/* body, two CRC octets, DEL */
			o.kind = crc ? AVLC_ACARS_BADCRC : AVLC_ACARS;
			if (!crc) {
				/* octets 0 .. n-2 lose their parity bit (outacars.c:223-226); anything at or beyond n-1 is only
				   reached in frames too short for a header and is taken as received, like the reference reads it */
#define AVLC_CH(i) ((uint8_t)((i) < n - 1 ? (t[i] & 0x7f) : t[i]))
				o.mode = AVLC_CH(0);
				for (int i = 0; i < 7; i++)
					o.reg[i] = AVLC_CH(1 + i);
				o.ack = AVLC_CH(8);
				if (o.ack == 0x15)
					o.ack = '!';
				o.label[0] = AVLC_CH(9);
				o.label[1] = AVLC_CH(10);
				if (o.label[1] == 0x7f)
					o.label[1] = 'd';
				o.bid = AVLC_CH(11);
				if (o.bid == 0)
					o.bid = ' ';
				o.bs = AVLC_CH(12);
				int k = 13;
				const int end = n - 4;	/* first octet behind the text: ETX/ETB, CRC, DEL follow */
				if (o.bs != 0x03) {
					if (o.mode <= 'Z' && o.bid <= '9') {
						while (o.nno < 4 && k < end)
							o.no[o.nno++] = AVLC_CH(k), k++;
						while (o.nfid < 6 && k < end)
							o.fid[o.nfid++] = AVLC_CH(k), k++;
					}
					o.txt_off = (uint16_t) (13 + k);
					if (k < end) {
						o.txt_len = (uint16_t) (end - k);
						k = end;
					}
				}
				o.be = AVLC_CH(k);
#undef AVLC_CH
			}
		} else
			o.kind = AVLC_OTHER;
	}
	*r = o;
}

VDL2_AVLC_FN void avlc_extract(const uint8_t * h, int l, Vdl2AvlcRec * r)
{
	uint32_t crc = 0;
	if (avlc_is_acars(h, l))
		for (int i = 0; i < l - 16 - 1; i++)
			crc = avlc_crc(crc, h[13 + i]);
	avlc_fields(h, l, crc, r);
}
#endif
