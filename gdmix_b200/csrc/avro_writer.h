// avro_writer.h -- host-side Avro binary encoding of the trainer's score records (no fastavro): the body of an
// object-container file for the `validation_result` schema of util/io_utils.py:367-375 -- uid (long),
// predictionScore (float), label ([null, float]), weight (float, when the dataset has the column),
// predictionScorePerCoordinate (float) -- in blocks of `records_per_block` records, each block
// <count varint> <size varint> <records> <16-byte sync>, exactly what batched_write_avro (:299-334) appends.
// The Python writer (gdmix_b200/io/avro.py) produces the same bytes one record at a time at ~120 k records/s.
#pragma once
#include <cmath>
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

// ---------------------------------------------------------------------------------------------------------
// Model files: Photon-ML BayesianLinearModelAvro records (models/schemas.py:3-51 as written by gen_one_avro_model,
// util/io_utils.py:102-160): modelId, modelClass, means = intercept first then every feature with
// |coefficient| > threshold as (name, term, value), variances aligned with means or null, lossFunction "".
// A Sink either counts bytes (first call, sizes the buffer) or writes them.
// ---------------------------------------------------------------------------------------------------------
struct Sink {
    uint8_t *p;      // nullptr: count only
    int64_t n = 0;
    void byte(uint8_t b) { if (p) p[n] = b; n++; }
    void lng(int64_t v)
    {
        uint64_t z = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);
        while (z >= 0x80) { byte((uint8_t)(z | 0x80)); z >>= 7; }
        byte((uint8_t)z);
    }
    void raw(const void *src, int64_t len) { if (p && len) memcpy(p + n, src, (size_t)len); n += len; }
    void str(const char *s, int64_t len) { lng(len); raw(s, len); }
    void dbl(double v) { raw(&v, 8); }
};

struct ModelTable {
    int64_t n_models;
    const char *id_chars; const int64_t *id_ptr;
    const char *model_class;
    const double *coef, *var; const int64_t *coef_ptr;       // per model: [intercept,] features...
    const int64_t *feat_idx;                                  // global feature index of every non-intercept coefficient
    int32_t has_intercept; double threshold;
    const char *intercept_name;
    const char *name_chars; const int64_t *name_ptr; const char *term_chars; const int64_t *term_ptr; int64_t n_features;
};

inline bool avro_one_model(const ModelTable &t, int64_t m, Sink &o)
{
    const int64_t c0 = t.coef_ptr[m], c1 = t.coef_ptr[m + 1];
    const int64_t hi = t.has_intercept ? 1 : 0;
    if (c1 - c0 < hi) return false;
    const int64_t f0 = c0 - m * hi;   // position of this model's first feature in feat_idx (intercepts are not listed)
    o.str(t.id_chars + t.id_ptr[m], t.id_ptr[m + 1] - t.id_ptr[m]);
    o.lng(1); o.str(t.model_class, (int64_t)strlen(t.model_class));
    int64_t kept = hi;
    for (int64_t j = c0 + hi; j < c1; j++) kept += (fabs(t.coef[j]) > t.threshold) ? 1 : 0;
    for (int pass = 0; pass < 2; pass++) {
        const double *v = pass == 0 ? t.coef : t.var;
        if (pass == 1) {
            if (!t.var) { o.lng(0); break; }   // variances: null
            o.lng(1);
        }
        if (kept) {
            o.lng(kept);
            if (hi) { o.str(t.intercept_name, (int64_t)strlen(t.intercept_name)); o.str("", 0); o.dbl(v[c0]); }
            for (int64_t j = c0 + hi; j < c1; j++) {
                if (!(fabs(t.coef[j]) > t.threshold)) continue;
                const int64_t g = t.feat_idx[f0 + (j - c0 - hi)];
                if (g < 0 || g >= t.n_features) return false;
                o.str(t.name_chars + t.name_ptr[g], t.name_ptr[g + 1] - t.name_ptr[g]);
                o.str(t.term_chars + t.term_ptr[g], t.term_ptr[g + 1] - t.term_ptr[g]);
                o.dbl(v[j]);
            }
        }
        o.lng(0);   // end of array
    }
    o.lng(1); o.lng(0);   // lossFunction: ""
    return true;
}

// blocks of `per_block` records; out == nullptr: returns the bytes needed.  -1: inconsistent input.
inline int64_t avro_model_blocks(const ModelTable &t, int32_t per_block, const uint8_t *sync, uint8_t *out)
{
    Sink o{out};
    for (int64_t b0 = 0; b0 < t.n_models; b0 += per_block) {
        const int64_t cnt = (t.n_models - b0 < per_block) ? t.n_models - b0 : per_block;
        Sink size_of{nullptr};
        for (int64_t m = b0; m < b0 + cnt; m++)
            if (!avro_one_model(t, m, size_of)) return -1;
        o.lng(cnt);
        o.lng(size_of.n);
        for (int64_t m = b0; m < b0 + cnt; m++) avro_one_model(t, m, o);
        o.raw(sync, 16);
    }
    return o.n;
}

}  // namespace gdmix_host
