// avro_writer.h -- host-side Avro binary encoding of the trainer's score records (no fastavro): the body of an
// object-container file for the `validation_result` schema of util/io_utils.py:367-375 -- uid (long),
// predictionScore (float), label ([null, float]), weight (float, when the dataset has the column),
// predictionScorePerCoordinate (float) -- in blocks of `records_per_block` records, each block
// <count varint> <size varint> <records> <16-byte sync>, exactly what batched_write_avro (:299-334) appends.
// The Python writer (gdmix_b200/io/avro.py) produces the same bytes one record at a time at ~120 k records/s.
#pragma once
#include <vector>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>

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

inline int long_bytes(int64_t v)
{
    uint64_t z = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);
    int n = 1;
    while (z >= 0x80) { n++; z >>= 7; }
    return n;
}

// Blocks are sized (only the uid varints vary), then written in place by all host threads.
inline int64_t avro_score_blocks(const int64_t *uid, const float *score, const float *label, const float *weight,
                                 const float *per_coord, int64_t n, int32_t per_block, const uint8_t *sync, uint8_t *out)
{
    if (n <= 0 || per_block <= 0) return 0;
    const int64_t nb = (n + per_block - 1) / per_block;
    const int64_t fixed = 4 + 1 + (label ? 4 : 0) + (weight ? 4 : 0) + (per_coord ? 4 : 0);
    std::vector<int64_t> start((size_t)nb + 1, 0), body((size_t)nb, 0);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nb; b++) {
        const int64_t b0 = b * per_block, cnt = (n - b0 < per_block) ? n - b0 : per_block;
        int64_t sz = cnt * fixed;
        for (int64_t i = b0; i < b0 + cnt; i++) sz += long_bytes(uid[i]);
        body[(size_t)b] = sz;
    }
    for (int64_t b = 0; b < nb; b++) {
        const int64_t b0 = b * per_block, cnt = (n - b0 < per_block) ? n - b0 : per_block;
        start[(size_t)b + 1] = start[(size_t)b] + long_bytes(cnt) + long_bytes(body[(size_t)b]) + body[(size_t)b] + 16;
    }
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nb; b++) {
        const int64_t b0 = b * per_block, cnt = (n - b0 < per_block) ? n - b0 : per_block;
        uint8_t *q = out + start[(size_t)b];
        q = put_long(q, cnt);
        q = put_long(q, body[(size_t)b]);
        for (int64_t i = b0; i < b0 + cnt; i++) {
            q = put_long(q, uid[i]);
            q = put_float(q, score[i]);
            if (label) { *q++ = 0x02; q = put_float(q, label[i]); } else *q++ = 0x00;   // union branch 1 (float) / 0 (null)
            if (weight) q = put_float(q, weight[i]);
            if (per_coord) q = put_float(q, per_coord[i]);
        }
        memcpy(q, sync, 16);
    }
    return start[(size_t)nb];
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
            const int64_t *fi = t.feat_idx + f0 - (c0 + hi);   // fi[j] = global feature of coefficient j
            for (int64_t j = c0 + hi; j < c1; j++) {
                // the (name, term) of a feature are four dependent cache misses in tables the size of the feature
                // space: fetch the pointers 16 coefficients ahead and the characters 8 ahead, so the misses overlap
                if (j + 16 < c1) {
                    const int64_t g2 = fi[j + 16];
                    if ((uint64_t)g2 < (uint64_t)t.n_features) { __builtin_prefetch(t.name_ptr + g2); __builtin_prefetch(t.term_ptr + g2); }
                }
                if (o.p && j + 8 < c1) {
                    const int64_t g1 = fi[j + 8];
                    if ((uint64_t)g1 < (uint64_t)t.n_features) {
                        __builtin_prefetch(t.name_chars + t.name_ptr[g1]);
                        __builtin_prefetch(t.term_chars + t.term_ptr[g1]);
                    }
                }
                if (!(fabs(t.coef[j]) > t.threshold)) continue;
                const int64_t g = fi[j];
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

// Blocks of `per_block` records, sized (avro_model_sizes: start[b] = first byte of block b, start[nb] = total; -1:
// inconsistent input) and then written (avro_model_write) by all host threads -- a block's bytes depend on nothing
// outside it.
inline int64_t avro_model_sizes(const ModelTable &t, int32_t per_block, std::vector<int64_t> &start, std::vector<int64_t> &body)
{
    const int64_t nb = (t.n_models + per_block - 1) / per_block;
    body.assign((size_t)nb, 0);
    start.assign((size_t)nb + 1, 0);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(| : bad)
    for (int64_t b = 0; b < nb; b++) {
        const int64_t b0 = b * per_block;
        const int64_t cnt = (t.n_models - b0 < per_block) ? t.n_models - b0 : per_block;
        Sink size_of{nullptr};
        for (int64_t m = b0; m < b0 + cnt; m++)
            if (!avro_one_model(t, m, size_of)) bad |= 1;
        body[(size_t)b] = size_of.n;
    }
    if (bad) return -1;
    for (int64_t b = 0; b < nb; b++) {
        const int64_t b0 = b * per_block;
        const int64_t cnt = (t.n_models - b0 < per_block) ? t.n_models - b0 : per_block;
        Sink hdr{nullptr};
        hdr.lng(cnt); hdr.lng(body[(size_t)b]);
        start[(size_t)b + 1] = start[(size_t)b] + hdr.n + body[(size_t)b] + 16;
    }
    return start[(size_t)nb];
}

inline void avro_model_write(const ModelTable &t, int32_t per_block, const uint8_t *sync, const std::vector<int64_t> &start,
                             const std::vector<int64_t> &body, uint8_t *out)
{
    const int64_t nb = (int64_t)body.size();
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < nb; b++) {
        const int64_t b0 = b * per_block;
        const int64_t cnt = (t.n_models - b0 < per_block) ? t.n_models - b0 : per_block;
        Sink o{out + start[(size_t)b]};
        o.lng(cnt);
        o.lng(body[(size_t)b]);
        for (int64_t m = b0; m < b0 + cnt; m++) avro_one_model(t, m, o);
        o.raw(sync, 16);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Reading model files back (warm starts, the predict action): the records of one container block of
// BayesianLinearModelAvro -> flat arrays.  Every (name, term) is looked up in the feature file's table; the
// intercept is ("(INTERCEPT)", "").  Count pass (null outputs) then fill pass, like the TFRecord readers.
// ---------------------------------------------------------------------------------------------------------
struct FeatureMap {
    std::unordered_map<std::string, int64_t> index;   // key = name + '\x01' + term
    std::string intercept;
};

struct ModelDecodeOut {   // null pointers: counting pass
    char *id_chars = nullptr; int64_t *id_ptr = nullptr;
    int64_t *mean_ptr = nullptr;      // [n + 1] into mean_feat / mean_val / var_val
    int64_t *mean_feat = nullptr;     // global feature index, -1 = the intercept
    double *mean_val = nullptr, *var_val = nullptr;
    uint8_t *has_var = nullptr;       // [n]
};
struct ModelDecodeSizes { int64_t n_models = 0, n_means = 0, id_bytes = 0; };

class ModelDecoder {
public:
    ModelDecoder(const FeatureMap &fm, std::string &err) : fm_(fm), err_(err) {}

    bool run(const uint8_t *buf, int64_t len, int64_t n_records, ModelDecodeSizes &sz, const ModelDecodeOut &o)
    {
        p_ = buf; e_ = buf + len;
        for (int64_t r = 0; r < n_records; r++)
            if (!record(sz, o)) return false;
        if (p_ != e_) return fail("trailing bytes after the block's records");
        if (o.id_ptr) o.id_ptr[sz.n_models] = sz.id_bytes;
        if (o.mean_ptr) o.mean_ptr[sz.n_models] = sz.n_means;
        return true;
    }

private:
    bool fail(const char *msg) { err_ = msg; return false; }
    bool lng(int64_t &v)
    {
        uint64_t z = 0;
        for (int shift = 0; shift < 70; shift += 7) {
            if (p_ >= e_) return false;
            const uint8_t b = *p_++;
            if (shift < 64) z |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) { v = (int64_t)(z >> 1) ^ -(int64_t)(z & 1); return true; }
        }
        return false;
    }
    bool str(const uint8_t *&s, int64_t &n)
    {
        if (!lng(n) || n < 0 || n > e_ - p_) return false;
        s = p_; p_ += n;
        return true;
    }
    bool dbl(double &v) { if (e_ - p_ < 8) return false; memcpy(&v, p_, 8); p_ += 8; return true; }
    bool opt_string()   // ["null", "string"]
    {
        int64_t br; const uint8_t *s; int64_t n;
        if (!lng(br)) return false;
        if (br == 0) return true;
        return br == 1 && str(s, n);
    }
    // one array of NameTermValueAvro; which = 0: means (defines the entry list), 1: variances (must align)
    bool ntv_array(int which, ModelDecodeSizes &sz, const ModelDecodeOut &o, int64_t first, int64_t &count)
    {
        count = 0;
        for (;;) {
            int64_t n;
            if (!lng(n)) return fail("malformed array block count");
            if (n == 0) break;
            if (n < 0) { int64_t bytes; n = -n; if (!lng(bytes)) return fail("malformed array block size"); }
            for (int64_t i = 0; i < n; i++) {
                const uint8_t *nm, *tm; int64_t nl, tl; double v;
                if (!str(nm, nl) || !str(tm, tl) || !dbl(v)) return fail("malformed name-term-value");
                int64_t feat;
                if (tl == 0 && (size_t)nl == fm_.intercept.size() && memcmp(nm, fm_.intercept.data(), nl) == 0) {
                    feat = -1;
                } else {
                    key_.assign((const char *)nm, (size_t)nl); key_.push_back('\x01'); key_.append((const char *)tm, (size_t)tl);
                    auto it = fm_.index.find(key_);
                    if (it == fm_.index.end()) { err_ = "feature (" + key_.substr(0, nl) + ", " + std::string((const char *)tm, (size_t)tl) + ") is not in the feature file"; return false; }
                    feat = it->second;
                }
                if (which == 0) {
                    if (o.mean_feat) { o.mean_feat[first + count] = feat; o.mean_val[first + count] = v; }
                } else {
                    if (o.mean_feat) {
                        if (o.mean_feat[first + count] != feat) return fail("variances are not aligned with means");
                        o.var_val[first + count] = v;
                    }
                }
                count++;
            }
        }
        return true;
    }
    bool record(ModelDecodeSizes &sz, const ModelDecodeOut &o)
    {
        const uint8_t *id; int64_t idl;
        if (!str(id, idl)) return fail("malformed modelId");
        if (o.id_chars) { o.id_ptr[sz.n_models] = sz.id_bytes; memcpy(o.id_chars + sz.id_bytes, id, (size_t)idl); }
        if (!opt_string()) return fail("malformed modelClass");
        const int64_t first = sz.n_means;
        if (o.mean_ptr) o.mean_ptr[sz.n_models] = first;
        int64_t nm = 0, nv = 0;
        if (!ntv_array(0, sz, o, first, nm)) return false;
        int64_t br;
        if (!lng(br)) return fail("malformed variances union");
        if (br == 1) {
            if (!ntv_array(1, sz, o, first, nv)) return false;
            if (nv != 0 && nv != nm) return fail("variances and means differ in length");
        } else if (br != 0) return fail("malformed variances union");
        if (o.has_var) o.has_var[sz.n_models] = nv ? 1 : 0;
        if (o.var_val && !nv) for (int64_t i = 0; i < nm; i++) o.var_val[first + i] = 0.0;
        if (!opt_string()) return fail("malformed lossFunction");
        sz.id_bytes += idl;
        sz.n_means += nm;
        sz.n_models++;
        return true;
    }

    const FeatureMap &fm_;
    std::string &err_;
    std::string key_;
    const uint8_t *p_ = nullptr, *e_ = nullptr;
};

}  // namespace gdmix_host
