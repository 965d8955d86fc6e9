/*
 * oracle/orc_avlc_api.h -- TEST INFRASTRUCTURE ONLY.
 * CPU checkers for SURVEY.md section 8(f) row f4: the fields the reference derives from a frame's bytes before it
 * formats anything -- out() (out.c:517-598: addresses through icaoaddr(), out.c:426-435; direction, command/response,
 * on-ground bit, link control byte, payload class) and outacars() (outacars.c:214-290: ACARS CRC, parity strip, mode,
 * registration, ack, label, block id, message number, flight id, text extent, end marker).
 *   oracle/ref/ref_out_harness.c  the reference's out.c + outacars.c + outxid.c + label.c + cJSON.c + crc.c compiled in
 *                                 place; returns the -J JSON line out() prints for a frame -> oracle/_ref/libvdl2outref.so
 *   oracle/port/vdl2_avlc_port.c  an independent plain-C restatement producing the binary record -> oracle/libvdl2avlcport.so
 * Only tests/ may load these libraries.  The device side of row f4 is not built yet.
 */
#ifndef ORC_AVLC_API_H
#define ORC_AVLC_API_H
#include <stdint.h>

enum orc_avlc_kind {
	ORC_AVLC_EMPTY = 0,	/* l <= 13: no information field (out.c:530) */
	ORC_AVLC_XID = 1,	/* hdata[10] == 0x82, l >= 14 (out.c:562) */
	ORC_AVLC_ACARS = 2,	/* ff ff 01 header, l >= 16, ACARS CRC good (out.c:566, outacars.c:222-230) */
	ORC_AVLC_ACARS_BADCRC = 3,
	ORC_AVLC_OTHER = 4	/* anything else with l > 13 ("unknown data", out.c:570) */
};

typedef struct {
	uint32_t faddr, taddr;	/* icaoaddr(&hdata[5]), icaoaddr(&hdata[1]): 3-bit type << 24 | 24-bit address */
	uint8_t fromair;	/* (faddr >> 24) == 1 */
	uint8_t rep;		/* (hdata[5] & 2) >> 1: 1 = response, 0 = command */
	uint8_t gnd;		/* (hdata[1] & 2) != 0: aircraft on ground */
	uint8_t lc;		/* hdata[9], the link control byte */
	uint8_t kind;		/* enum orc_avlc_kind */
	/* ACARS only (kind 2), parity bits stripped: */
	uint8_t mode, ack, bid, bs, be;	/* ack: 0x15 -> '!'; bid: 0 -> ' ' (outacars.c:243-261) */
	uint8_t label[2];	/* label[1]: 0x7f -> 'd' */
	uint8_t reg[7];		/* the 7 registration characters as sent (fixreg()'s formatting stays on the host) */
	uint8_t nno, nfid;	/* characters in no[] / fid[] */
	uint8_t no[4], fid[6];
	uint8_t pad;
	uint16_t txt_off, txt_len;	/* message text = hdata[txt_off .. txt_off + txt_len), each byte & 0x7f */
	uint16_t info_off, info_len;	/* information field of any kind: hdata[10 .. l - 3) */
} orc_avlc;			/* 48 B */

#ifdef __cplusplus
extern "C" {
#endif
/* port: the record of one frame (hdata, l as handed to out()); hdata is not modified */
void orc_avlc_extract(const uint8_t * hdata, int l, orc_avlc * rec);
/* reference: the JSON line out() prints with -J -G -E for this frame (empty string if it prints none); returns its length */
int orc_out_json(const uint8_t * hdata, int l, int chn, int Fr, float ppm, double t, char *buf, int cap);
/* reference: the TEXT out() prints at the default verbosity with -G -E -U for this frame; returns its length */
int orc_out_text(const uint8_t * hdata, int l, int chn, int Fr, float ppm, double t, char *buf, int cap);
#ifdef __cplusplus
}
#endif
#endif
