/* TEST INFRASTRUCTURE ONLY: plain-C restatement of the field extraction the reference performs on a frame before it formats
   text or JSON.  Follows out.c:517-570 (out), out.c:426-435 (icaoaddr), outacars.c:214-290
   (outacars) and crc.h:3 (update_crc, reflected CRC-16/CCITT, polynomial 0x8408, started at 0).  Independent of the
   reference's sources: the CRC is computed bit by bit, the address bits are gathered directly. */
#include <string.h>
#include "../orc_avlc_api.h"

/* An AVLC address is 4 octets; octet 0 carries 6 address bits in bits 2..7, octets 1..3 carry 7 in bits 1..7, each group
   sent least significant bit first, so the value is the concatenation of the groups with their bits reversed. */
static uint32_t avlc_address(const uint8_t * a)
{
	uint32_t v = 0;
	for (int b = 2; b <= 7; b++)
		v = (v << 1) | ((a[0] >> b) & 1u);
	for (int o = 1; o < 4; o++)
		for (int b = 1; b <= 7; b++)
			v = (v << 1) | ((a[o] >> b) & 1u);
	return v;
}

static uint16_t crc_step(uint16_t crc, uint8_t c)
{
	crc ^= c;
	for (int i = 0; i < 8; i++)
		crc = (crc & 1u) ? (uint16_t) ((crc >> 1) ^ 0x8408u) : (uint16_t) (crc >> 1);
	return crc;
}

void orc_avlc_extract(const uint8_t * hdata, int l, orc_avlc * r)
{
	memset(r, 0, sizeof *r);
	r->faddr = avlc_address(hdata + 5);
	r->taddr = avlc_address(hdata + 1);
	r->fromair = (r->faddr >> 24) == 1;
	r->rep = (hdata[5] >> 1) & 1;
	r->gnd = (hdata[1] >> 1) & 1;
	r->lc = hdata[9];
	if (l <= 13) {
		r->kind = ORC_AVLC_EMPTY;
		return;
	}
	r->info_off = 10;
	r->info_len = (uint16_t) (l - 13);
	if (l >= 14 && hdata[10] == 0x82) {
		r->kind = ORC_AVLC_XID;
		return;
	}
	if (!(l >= 16 && hdata[10] == 0xff && hdata[11] == 0xff && hdata[12] == 0x01)) {
		r->kind = ORC_AVLC_OTHER;
		return;
	}
	const uint8_t *t = hdata + 13;
	const int len = l - 16;	/* body + 2 CRC octets + DEL */
	uint16_t crc = 0;
	for (int i = 0; i < len - 1; i++)
		crc = crc_step(crc, t[i]);
	if (crc) {
		r->kind = ORC_AVLC_ACARS_BADCRC;
		return;
	}
	r->kind = ORC_AVLC_ACARS;
	/* the reference strips the parity bit of t[0 .. len-2] in place while it computes the CRC (outacars.c:223-226); what it
	   reads beyond that (only in frames too short to hold a header) is read as received */
#define CH(i) ((uint8_t)((i) < len - 1 ? (t[i] & 0x7f) : t[i]))
	int k = 0;
	r->mode = CH(k);
	k++;
	for (int i = 0; i < 7; i++, k++)
		r->reg[i] = CH(k);
	r->ack = CH(k);
	k++;
	if (r->ack == 0x15)
		r->ack = '!';
	r->label[0] = CH(k);
	k++;
	r->label[1] = CH(k);
	k++;
	if (r->label[1] == 0x7f)
		r->label[1] = 'd';
	r->bid = CH(k);
	k++;
	if (r->bid == 0)
		r->bid = ' ';
	r->bs = CH(k);
	k++;
	if (r->bs != 0x03) {
		if (r->mode <= 'Z' && r->bid <= '9') {
			for (; r->nno < 4 && k < len - 4; k++)
				r->no[r->nno++] = CH(k);
			for (; r->nfid < 6 && k < len - 4; k++)
				r->fid[r->nfid++] = CH(k);
		}
		r->txt_off = (uint16_t) (13 + k);
		if (k < len - 4) {
			r->txt_len = (uint16_t) (len - 4 - k);
			k = len - 4;
		}
	}
	r->be = CH(k);
#undef CH
}
