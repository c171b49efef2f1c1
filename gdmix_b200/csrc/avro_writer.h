// avro_writer.h -- host-side Avro binary encoding of the trainer's score records (no fastavro): the body of an
// object-container file for the `validation_result` schema of util/io_utils.py:367-375 -- uid (long),
// predictionScore (float), label ([null, float]), weight (float, when the dataset has the column),
// predictionScorePerCoordinate (float) -- in blocks of `records_per_block` records, each block
// <count varint> <size varint> <records> <16-byte sync>, exactly what batched_write_avro (:299-334) appends.
// The Python writer (gdmix_b200/io/avro.py) produces the same bytes one record at a time at ~120 k records/s.
#pragma once
#include <cstdint>
#include <cstring>

namespace gdmix_host {

inline uint8_t *put_long(uint8_t *p, int64_t v)
{
    uint64_t z = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);   // zig-zag
    while (z >= 0x80) { *p++ = (uint8_t)(z | 0x80); z >>= 7; }
    *p++ = (uint8_t)z;
    return p;
}
inline uint8_t *put_float(uint8_t *p, float v) { memcpy(p, &v, 4); return p + 4; }

// upper bound on the bytes avro_score_blocks writes
inline int64_t avro_score_blocks_bound(int64_t n, int32_t per_block)
{
    const int64_t blocks = per_block > 0 ? (n + per_block - 1) / per_block : 0;
    return n * (10 + 4 + 1 + 4 + 4 + 4) + blocks * (10 + 10 + 16);
}

inline int64_t avro_score_blocks(const int64_t *uid, const float *score, const float *label, const float *weight,
                                 const float *per_coord, int64_t n, int32_t per_block, const uint8_t *sync, uint8_t *out)
{
    uint8_t *p = out;
    uint8_t tmp[24];
    for (int64_t b0 = 0; b0 < n; b0 += per_block) {
        const int64_t cnt = (n - b0 < per_block) ? n - b0 : per_block;
        // records first (into place after a gap for the two varints), then the header is moved in front
        uint8_t *body = p + 20;
        uint8_t *q = body;
        for (int64_t i = b0; i < b0 + cnt; i++) {
            q = put_long(q, uid[i]);
            q = put_float(q, score[i]);
            if (label) { *q++ = 0x02; q = put_float(q, label[i]); } else *q++ = 0x00;   // union branch 1 (float) / 0 (null)
            if (weight) q = put_float(q, weight[i]);
            if (per_coord) q = put_float(q, per_coord[i]);
        }
        const int64_t size = q - body;
        uint8_t *h = put_long(tmp, cnt);
        h = put_long(h, size);
        const int64_t hl = h - tmp;
        memcpy(p, tmp, hl);
        memmove(p + hl, body, size);
        p += hl + size;
        memcpy(p, sync, 16);
        p += 16;
    }
    return p - out;
}

}  // namespace gdmix_host
