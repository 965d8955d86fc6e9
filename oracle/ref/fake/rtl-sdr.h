/* TEST INFRASTRUCTURE: the 12 librtlsdr entry points rtl.c calls (rtl.c:52-302), backed by a file. */
#ifndef FAKE_RTL_SDR_H
#define FAKE_RTL_SDR_H
#include <stdint.h>
typedef struct rtlsdr_dev rtlsdr_dev_t;
typedef void (*rtlsdr_read_async_cb_t) (unsigned char *buf, uint32_t len, void *ctx);
uint32_t rtlsdr_get_device_count(void);
int rtlsdr_get_device_usb_strings(uint32_t index, char *manufact, char *product, char *serial);
const char *rtlsdr_get_device_name(uint32_t index);
int rtlsdr_open(rtlsdr_dev_t ** dev, uint32_t index);
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t * dev, int manual);
int rtlsdr_set_tuner_gain(rtlsdr_dev_t * dev, int gain);
int rtlsdr_get_tuner_gains(rtlsdr_dev_t * dev, int *gains);
int rtlsdr_set_freq_correction(rtlsdr_dev_t * dev, int ppm);
int rtlsdr_set_center_freq(rtlsdr_dev_t * dev, uint32_t freq);
int rtlsdr_set_sample_rate(rtlsdr_dev_t * dev, uint32_t rate);
int rtlsdr_reset_buffer(rtlsdr_dev_t * dev);
int rtlsdr_read_async(rtlsdr_dev_t * dev, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t buf_num, uint32_t buf_len);
#endif
