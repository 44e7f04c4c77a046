// celt_plc.cuh — packet-loss concealment (celt_decode_lost, opus-fix/celt/celt_decoder.c:415-711).
//
// SURVEY.md §8(f) rank 1 ("next" row).  NOT IMPLEMENTED YET: a lost CELT frame on a stream that has already
// decoded audio returns OPUS_UNIMPLEMENTED and leaves the state untouched (no CPU fallback by design).
#pragma once
#include "celt_decoder.cuh"

namespace cb {

template <class TM>
CB_DEV int celt_decode_lost_frame(TM, CbDecState *, SynthScratch &, int *const *, int) {
    return OPUS_UNIMPLEMENTED_;
}

}  // namespace cb
