// TEMPORARY stubs for entry points under construction (removed as each part lands).
#include "common.cuh"
using sdr::fail;
#define NI return fail(SDR_E_STATE, "%s: not implemented yet", __func__)
extern "C" {
int sdr_chan_new(const sdr_chan_config *, const float *, const uint32_t *, int, sdr_chan **) { NI; }
void sdr_chan_free(sdr_chan *) {}
int sdr_chan_reset(sdr_chan *) { NI; }
long sdr_chan_process(sdr_chan *, const uint8_t *, size_t, float *, float *, size_t) { NI; }
long sdr_chan_process_dev(sdr_chan *, const uint8_t *, size_t, float *, float *, size_t) { NI; }
int sdr_chan_sync(sdr_chan *) { NI; }
int sdr_chan_last_timing(const sdr_chan *, float *, uint32_t *) { NI; }
int sdr_comm_unique_id(uint8_t *) { NI; }
int sdr_comm_init(int, int, int, const uint8_t *, sdr_comm **) { NI; }
int sdr_comm_bcast_u8(sdr_comm *, uint8_t *, size_t, int) { NI; }
int sdr_comm_chan_wait(sdr_comm *, sdr_chan *) { NI; }
int sdr_comm_wait_chan(sdr_comm *, sdr_chan *) { NI; }
int sdr_comm_sync(sdr_comm *) { NI; }
void sdr_comm_free(sdr_comm *) {}
}
