/*
 * vdl2_multi.cu -- several B200s behind ONE handle of the C ABI (include/vdl2gpu.h, vdl2_multi_*).
 *
 * The path shards with no exchange step: input stream s, with the channels demodulated from it, lives
 * on device s mod N; every device runs the same kernel on its own streams; the host merges the completed blocks in the
 * order one device would have produced them and feeds the single consumer (the reference's blk_thread queue,
 * vdlm2.c:189-206).  One host thread drives all devices: uploads and launches are asynchronous (vdl2_submit_host), the
 * streams of a device are picked out of the caller's buffer by a strided 2-D copy (stream pitch N x pitch), and only then
 * does the call wait for every device.  No NCCL, no peer traffic: NVLink is idle by design.
 */
#include <algorithm>
#include <string>
#include <vector>
#include <stdio.h>
#include <string.h>
#include "../../include/vdl2gpu.h"

struct vdl2multi {
	std::vector < vdl2gpu_t * >dev;	/* one handle per device that owns at least one stream */
	std::vector < int >first;	/* global index of the device's first stream (= its position in the device list) */
	int ndev, nstreams, bytes_per_sample;
	std::string err;
	std::vector < vdl2_block_t > tmp;
};

static thread_local std::string g_multi_error;

static int mfail(vdl2multi * m, const std::string & msg)
{
	fprintf(stderr, "vdl2gpu: %s\n", msg.c_str());
	if (m)
		m->err = msg;
	else
		g_multi_error = msg;
	return 1;
}

extern "C" const char *vdl2_multi_last_error(const vdl2multi_t * m)
{
	return m ? m->err.c_str() : g_multi_error.c_str();
}

extern "C" int vdl2_multi_destroy(vdl2multi_t * m)
{
	if (!m)
		return 0;
	for (auto h:m->dev)
		vdl2_destroy(h);
	delete m;
	return 0;
}

extern "C" int vdl2_multi_create(const vdl2_config_t * cfg, const vdl2_chan_param_t * chans, const int *devices, int ndev, vdl2multi_t ** out)
{
	if (!cfg || !chans || !devices || !out || ndev <= 0)
		return mfail(NULL, "vdl2_multi_create: null argument or no device");
	*out = NULL;
	if (cfg->nch <= 0 || cfg->ch_per_stream <= 0 || cfg->nch % cfg->ch_per_stream)
		return mfail(NULL, "vdl2_multi_create: nch must be a positive multiple of ch_per_stream");
	vdl2multi *m = new vdl2multi();
	m->ndev = ndev;
	m->nstreams = cfg->nch / cfg->ch_per_stream;
	static const int bps[5] = { 2, 2, 4, 8, 4 };
	m->bytes_per_sample = (cfg->format >= 0 && cfg->format < 5) ? bps[cfg->format] : 0;
	const int cps = cfg->ch_per_stream;
	for (int d = 0; d < ndev && d < m->nstreams; d++) {
		std::vector < vdl2_chan_param_t > sub;
		for (int s = d; s < m->nstreams; s += ndev)
			for (int k = 0; k < cps; k++)
				sub.push_back(chans[s * cps + k]);
		vdl2_config_t c = *cfg;
		c.nch = (int)sub.size();
		c.device = devices[d];
		vdl2gpu_t *h = NULL;
		if (vdl2_create(&c, sub.data(), &h)) {
			const std::string why = vdl2_last_error(NULL);
			vdl2_multi_destroy(m);
			return mfail(NULL, "vdl2_multi_create: device " + std::to_string(devices[d]) + ": " + why);
		}
		m->dev.push_back(h);
		m->first.push_back(d);
	}
	*out = m;
	return 0;
}

extern "C" int vdl2_multi_ndev(const vdl2multi_t * m)
{
	return m ? (int)m->dev.size() : 0;
}

extern "C" vdl2gpu_t *vdl2_multi_handle(vdl2multi_t * m, int i)
{
	return (m && i >= 0 && i < (int)m->dev.size())? m->dev[i] : NULL;
}

extern "C" int vdl2_multi_process_host(vdl2multi_t * m, const void *iq, size_t nsamples, size_t pitch_bytes)
{
	if (!m || !iq)
		return mfail(m, "vdl2_multi_process_host: null argument");
	if (m->nstreams > 1 && pitch_bytes < nsamples * (size_t) m->bytes_per_sample)
		return mfail(m, "vdl2_multi_process_host: pitch smaller than a stream");
	const size_t n = m->dev.size();
	/* enqueue everywhere first: device d takes streams d, d + N, ... = a 2-D copy with stream pitch N x pitch */
	for (size_t d = 0; d < n; d++)
		if (vdl2_submit_host(m->dev[d], (const uint8_t *)iq + (size_t) m->first[d] * pitch_bytes, nsamples, pitch_bytes * (size_t) m->ndev))
			return mfail(m, std::string("device ") + std::to_string(d) + ": " + vdl2_last_error(m->dev[d]));
	for (size_t d = 0; d < n; d++)
		if (vdl2_sync(m->dev[d]))
			return mfail(m, std::string("device ") + std::to_string(d) + ": " + vdl2_last_error(m->dev[d]));
	return 0;
}

extern "C" int vdl2_multi_drain_blocks(vdl2multi_t * m, vdl2_block_t * out, int max, int *n_out)
{
	if (!m || !n_out)
		return mfail(m, "vdl2_multi_drain_blocks: null argument");
	*n_out = 0;
	int total = 0;
	for (size_t d = 0; d < m->dev.size(); d++) {
		int n = 0;
		if (vdl2_drain_blocks(m->dev[d], out ? out + total : NULL, max - total, &n))
			return mfail(m, std::string("device ") + std::to_string(d) + ": " + vdl2_last_error(m->dev[d]));
		total += n;
	}
	/* the order one device would have produced: oldest trigger first, then channel (every per-device part is already sorted) */
	if (total > 1 && m->dev.size() > 1) {
		std::vector < int >idx(total);
		for (int i = 0; i < total; i++)
			idx[i] = i;
		std::stable_sort(idx.begin(), idx.end(),[&](int a, int b) {
				 return out[a].sync_dump != out[b].sync_dump ? out[a].sync_dump < out[b].sync_dump : out[a].chn < out[b].chn;}
		);
		m->tmp.assign(out, out + total);
		for (int i = 0; i < total; i++)
			out[i] = m->tmp[idx[i]];
	}
	*n_out = total;
	return 0;
}
