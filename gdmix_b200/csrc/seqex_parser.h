// seqex_parser.h -- host-side reader of entity-grouped TFRecord files (no TensorFlow): the framing of a TFRecord
// file and the protobuf wire format of tf.train.SequenceExample, decoded straight into the flat arrays the
// random-effect ingest builds (gdmix_b200/ingest.py: read_entity_grouped), which is what
// per_entity_grouped_input_fn + prepare_jobs hand to the consumers in the reference
// (gdmix-trainer/src/gdmix/io/input_data_pipeline.py:244-273, models/custom/scipy/job_consumers.py:161-258).
//
// One record = one entity:
//   context        entity id (int64_list | bytes_list, one value), and one variable-length list per sample column
//                  (uid int64, label int64|float, offset float, weight float)
//   feature_lists  <bag>_indices: one int64_list Feature per sample, <bag>_values: one float_list Feature per sample
//
// Two passes over the same buffer: count (entities, samples, non-zeros, id characters), then fill caller-allocated
// arrays.  Anything malformed is an error with a message, never a silent skip.
#pragma once
#include <vector>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/gdmix_b200.h"

namespace gdmix_host {

struct Span {
    const uint8_t *p, *e;
    bool empty() const { return p >= e; }
};

inline bool rd_varint(Span &s, uint64_t &v)
{
    if (s.e - s.p >= 3) {   // one to three bytes (values below 2^21: feature indices, lengths, keys) without the loop
        const uint8_t *p = s.p;
        uint64_t b = p[0], r = b & 0x7f;
        if (!(b & 0x80)) { s.p = p + 1; v = r; return true; }
        b = p[1]; r |= (b & 0x7f) << 7;
        if (!(b & 0x80)) { s.p = p + 2; v = r; return true; }
        b = p[2]; r |= (b & 0x7f) << 14;
        if (!(b & 0x80)) { s.p = p + 3; v = r; return true; }
    }
    v = 0;
    for (int shift = 0; shift < 70 && s.p < s.e; shift += 7) {
        const uint8_t b = *s.p++;
        if (shift < 64) v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return true;
    }
    return false;
}

// next field of a message: number, wire type, and for length-delimited fields its payload
inline bool rd_field(Span &s, uint32_t &fno, uint32_t &wt, uint64_t &scalar, Span &sub)
{
    uint64_t key;
    if (!rd_varint(s, key)) return false;
    fno = (uint32_t)(key >> 3); wt = (uint32_t)(key & 7);
    switch (wt) {
    case 0: return rd_varint(s, scalar);
    case 1: if (s.e - s.p < 8) return false; memcpy(&scalar, s.p, 8); s.p += 8; return true;
    case 5: { if (s.e - s.p < 4) return false; uint32_t t; memcpy(&t, s.p, 4); scalar = t; s.p += 4; return true; }
    case 2: {
        uint64_t ln;
        if (!rd_varint(s, ln) || ln > (uint64_t)(s.e - s.p)) return false;
        sub.p = s.p; sub.e = s.p + ln; s.p += ln;
        return true;
    }
    default: return false;
    }
}

enum Kind { kNone = 0, kBytes = 1, kFloat = 2, kInt64 = 3 };

// A tf.train.Feature: which list it holds and where that list's payload is.
inline bool feature_kind(Span f, Kind &kind, Span &list)
{
    kind = kNone; list = Span{nullptr, nullptr};
    uint32_t fno, wt; uint64_t sc; Span sub;
    while (!f.empty()) {
        if (!rd_field(f, fno, wt, sc, sub)) return false;
        if (wt == 2 && fno >= 1 && fno <= 3) { kind = (Kind)fno; list = sub; return true; }
    }
    return true;   // an empty Feature
}

// FloatList / Int64List payloads: repeated field 1, packed (wire type 2) or one element per key
template <class F>
inline bool each_float(Span list, F &&f)
{
    uint32_t fno, wt; uint64_t sc; Span sub;
    while (!list.empty()) {
        if (!rd_field(list, fno, wt, sc, sub)) return false;
        if (fno != 1) continue;
        if (wt == 2) {
            if ((sub.e - sub.p) % 4) return false;
            for (const uint8_t *q = sub.p; q < sub.e; q += 4) { float v; memcpy(&v, q, 4); f(v); }
        } else if (wt == 5) {
            const uint32_t bits = (uint32_t)sc; float v; memcpy(&v, &bits, 4); f(v);
        } else return false;
    }
    return true;
}
template <class F>
inline bool each_int64(Span list, F &&f)
{
    uint32_t fno, wt; uint64_t sc; Span sub;
    while (!list.empty()) {
        if (!rd_field(list, fno, wt, sc, sub)) return false;
        if (fno != 1) continue;
        if (wt == 2) {
            while (!sub.empty()) { uint64_t v; if (!rd_varint(sub, v)) return false; f((int64_t)v); }
        } else if (wt == 0) {
            f((int64_t)sc);
        } else return false;
    }
    return true;
}

// number of varints in a packed payload = bytes without the continuation bit (a loop the compiler vectorises);
// false when the payload ends inside a varint
inline bool count_varints(Span sub, int64_t &n)
{
    if (sub.empty()) return true;
    int64_t c = 0;
    for (const uint8_t *q = sub.p; q < sub.e; q++) c += (*q & 0x80) == 0;
    n += c;
    return (sub.e[-1] & 0x80) == 0;
}
// An Int64List written the usual way -- ONE packed field 1 and nothing else
// A FloatList written the usual way -- ONE packed field 1 and nothing else: its payload (copied, not walked)
inline bool single_packed(Span list, Span &payload)
{
    uint32_t fno, wt; uint64_t sc; Span sub;
    if (list.empty() || !rd_field(list, fno, wt, sc, sub)) return false;
    if (fno != 1 || wt != 2 || !list.empty()) return false;
    payload = sub;
    return true;
}

// number of elements of a list of the given kind (-1: malformed)
inline int64_t count_elems(Kind k, Span list)
{
    int64_t n = 0;
    if (k == kFloat) return each_float(list, [&](float) { n++; }) ? n : -1;
    if (k == kInt64) return each_int64(list, [&](int64_t) { n++; }) ? n : -1;
    if (k == kBytes) {
        uint32_t fno, wt; uint64_t sc; Span sub;
        while (!list.empty()) {
            if (!rd_field(list, fno, wt, sc, sub)) return -1;
            if (fno == 1) n++;
        }
    }
    return n;
}

struct SeqexOut {   // null pointers: counting pass
    int64_t *ent_rows = nullptr, *row_len = nullptr, *gcol = nullptr, *uid = nullptr, *id_ptr = nullptr;
    float *val = nullptr, *label = nullptr, *offset = nullptr, *weight = nullptr;
    char *id_chars = nullptr;
    // fused entity-local indexing (instead of gcol): local16[nnz] = rank of every index among its entity's distinct
    // indices, d_e[E] their number, uniq_scratch[nnz] the sorted distinct indices of entity e at its first non-zero
    uint16_t *local16 = nullptr;
    int64_t *d_e = nullptr, *uniq_scratch = nullptr;
};

class SeqexReader {
public:
    SeqexReader(const gdmix_seqex_spec &spec, std::string &err) : spec_(spec), err_(err) {}

    // All records of one uncompressed TFRecord file image.  `o` all-null = count only.
    bool run(const uint8_t *buf, int64_t len, gdmix_seqex_sizes &sz, const SeqexOut &o)
    {
        memset(&sz, 0, sizeof(sz));
        sz.all_labelled = 1;
        sz.min_index = INT64_MAX; sz.max_index = INT64_MIN;
        const uint8_t *p = buf, *end = buf + len;
        int64_t rec = 0;
        while (p < end) {
            if (end - p < 12) return fail("truncated TFRecord header at byte %lld", (long long)(p - buf));
            uint64_t n; memcpy(&n, p, 8);
            if (n > (uint64_t)(end - p - 12) || (uint64_t)(end - p - 12) - n < 4)
                return fail("truncated TFRecord payload at byte %lld", (long long)(p - buf));
            Span payload{p + 12, p + 12 + n};
            if (!record(payload, sz, o, rec)) return false;
            p += 12 + n + 4;
            rec++;
        }
        if (o.id_ptr) o.id_ptr[sz.n_entities] = sz.id_bytes;
        return true;
    }

private:
    bool fail(const char *fmt, long long a = 0, long long b = 0, long long c = 0)
    {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), fmt, a, b, c);
        err_ = tmp;
        return false;
    }
    static bool name_is(Span key, const char *name)
    {
        if (!name) return false;
        const size_t n = strlen(name);
        return (size_t)(key.e - key.p) == n && memcmp(key.p, name, n) == 0;
    }

    bool record(Span rec, gdmix_seqex_sizes &sz, const SeqexOut &o, int64_t rec_no)
    {
        Span context{nullptr, nullptr}, lists{nullptr, nullptr};
        uint32_t fno, wt; uint64_t sc; Span sub;
        while (!rec.empty()) {
            if (!rd_field(rec, fno, wt, sc, sub)) return fail("record %lld: malformed SequenceExample", rec_no);
            if (wt == 2 && fno == 1) context = sub;
            else if (wt == 2 && fno == 2) lists = sub;
        }
        // ---- context: entity id and the per-sample columns
        Span f_entity{nullptr, nullptr}, f_uid = f_entity, f_label = f_entity, f_off = f_entity, f_w = f_entity;
        bool has_entity = false, has_uid = false, has_label = false, has_off = false, has_w = false;
        Span c = context;
        while (!c.empty()) {
            if (!rd_field(c, fno, wt, sc, sub)) return fail("record %lld: malformed context", rec_no);
            if (fno != 1 || wt != 2) continue;
            Span key{nullptr, nullptr}, val{nullptr, nullptr};
            Span entry = sub;
            while (!entry.empty()) {
                uint32_t f2, w2; uint64_t s2; Span sub2;
                if (!rd_field(entry, f2, w2, s2, sub2)) return fail("record %lld: malformed context entry", rec_no);
                if (w2 == 2 && f2 == 1) key = sub2;
                else if (w2 == 2 && f2 == 2) val = sub2;
            }
            if (name_is(key, spec_.entity)) { f_entity = val; has_entity = true; }
            else if (name_is(key, spec_.uid)) { f_uid = val; has_uid = true; }
            else if (name_is(key, spec_.label)) { f_label = val; has_label = true; }
            else if (name_is(key, spec_.offset)) { f_off = val; has_off = true; }
            else if (name_is(key, spec_.weight)) { f_w = val; has_w = true; }
        }
        if (!has_entity) return fail("record %lld without the entity column", rec_no);
        if (!has_uid) return fail("record %lld without the uid column", rec_no);
        const int64_t e = sz.n_entities;
        // entity id -> its decimal / utf-8 string
        {
            Kind k; Span l;
            if (!feature_kind(f_entity, k, l) || k == kNone) return fail("record %lld: empty entity id", rec_no);
            if (k == kFloat) return fail("record %lld: float entity ids take the Python reader", rec_no);
            if (o.id_ptr) o.id_ptr[e] = sz.id_bytes;
            if (k == kInt64) {
                bool first = true; int64_t idv = 0;
                if (!each_int64(l, [&](int64_t v) { if (first) { idv = v; first = false; } }) || first)
                    return fail("record %lld: empty entity id", rec_no);
                char tmp[32];
                const int n = snprintf(tmp, sizeof(tmp), "%lld", (long long)idv);
                if (o.id_chars) memcpy(o.id_chars + sz.id_bytes, tmp, n);
                sz.id_bytes += n;
            } else {
                uint32_t f2, w2; uint64_t s2; Span sub2; bool got = false;
                while (!l.empty() && !got) {
                    if (!rd_field(l, f2, w2, s2, sub2)) return fail("record %lld: malformed entity id", rec_no);
                    if (f2 == 1 && w2 == 2) {
                        if (o.id_chars) memcpy(o.id_chars + sz.id_bytes, sub2.p, sub2.e - sub2.p);
                        sz.id_bytes += sub2.e - sub2.p;
                        got = true;
                    }
                }
                if (!got) return fail("record %lld: empty entity id", rec_no);
            }
        }
        // uid defines the number of samples
        const int64_t row0 = sz.n_rows;
        int64_t n = 0;
        {
            Kind k; Span l;
            if (!feature_kind(f_uid, k, l) || (k != kInt64 && k != kNone)) return fail("record %lld: uid must be an int64 list", rec_no);
            if (!each_int64(l, [&](int64_t v) { if (o.uid) o.uid[row0 + n] = v; n++; }))
                return fail("record %lld: malformed uid list", rec_no);
        }
        if (o.ent_rows) o.ent_rows[e] = n;
        auto column = [&](Span f, bool present, float *dst, float dflt, const char *what) -> bool {
            if (!present) {
                if (dst) for (int64_t i = 0; i < n; i++) dst[row0 + i] = dflt;
                return true;
            }
            Kind k; Span l;
            if (!feature_kind(f, k, l)) return fail("record %lld: malformed column", rec_no);
            int64_t m = 0;
            bool ok = true;
            if (k == kFloat) ok = each_float(l, [&](float v) { if (dst && m < n) dst[row0 + m] = v; m++; });
            else if (k == kInt64) ok = each_int64(l, [&](int64_t v) { if (dst && m < n) dst[row0 + m] = (float)v; m++; });
            else if (k == kBytes) return fail("record %lld: a bytes list where numbers are expected", rec_no);
            if (!ok) return fail("record %lld: malformed numeric list", rec_no);
            if (m != n) { (void)what; return fail("record %lld: a column has %lld values for %lld samples", rec_no, m, n); }
            return true;
        };
        if (has_label) { if (!column(f_label, true, o.label, 0.0f, "label")) return false; }
        else sz.all_labelled = 0;
        if (!column(f_off, has_off, o.offset, 0.0f, "offset")) return false;
        if (!column(f_w, has_w, o.weight, 1.0f, "weight")) return false;
        if (has_w) sz.saw_weight = 1;
        // ---- feature lists: <bag>_indices and <bag>_values, one Feature per sample
        const int64_t ent_q0 = sz.nnz;
        ent_g_.clear();
        Span l_idx{nullptr, nullptr}, l_val{nullptr, nullptr};
        bool has_idx = false, has_val = false;
        Span fl = lists;
        while (!fl.empty()) {
            if (!rd_field(fl, fno, wt, sc, sub)) return fail("record %lld: malformed feature_lists", rec_no);
            if (fno != 1 || wt != 2) continue;
            Span key{nullptr, nullptr}, val{nullptr, nullptr};
            Span entry = sub;
            while (!entry.empty()) {
                uint32_t f2, w2; uint64_t s2; Span sub2;
                if (!rd_field(entry, f2, w2, s2, sub2)) return fail("record %lld: malformed feature_lists entry", rec_no);
                if (w2 == 2 && f2 == 1) key = sub2;
                else if (w2 == 2 && f2 == 2) val = sub2;
            }
            if (name_is(key, spec_.bag_indices)) { l_idx = val; has_idx = true; }
            else if (name_is(key, spec_.bag_values)) { l_val = val; has_val = true; }
        }
        (void)has_idx; (void)has_val;
        // walk both FeatureLists in step
        int64_t si = 0, sv = 0;
        Span a = l_idx, b = l_val;
        for (;;) {
            Span fa{nullptr, nullptr}, fb{nullptr, nullptr};
            bool ga = false, gb = false;
            while (!a.empty() && !ga) {
                if (!rd_field(a, fno, wt, sc, sub)) return fail("record %lld: malformed index list", rec_no);
                if (fno == 1 && wt == 2) { fa = sub; ga = true; }
            }
            while (!b.empty() && !gb) {
                if (!rd_field(b, fno, wt, sc, sub)) return fail("record %lld: malformed value list", rec_no);
                if (fno == 1 && wt == 2) { fb = sub; gb = true; }
            }
            if (!ga && !gb) break;
            if (ga) si++;
            if (gb) sv++;
            if (ga != gb) continue;   // counted; the mismatch is reported below
            Kind ka, kb; Span la, lb;
            if (!feature_kind(fa, ka, la) || !feature_kind(fb, kb, lb)) return fail("record %lld: malformed Feature", rec_no);
            // an EMPTY list of another kind is still an empty sample (the encoder has no type to go by)
            if (ka != kInt64 && ka != kNone) { if (count_elems(ka, la) != 0) return fail("record %lld: feature indices must be int64 lists", rec_no); ka = kNone; }
            if (kb != kFloat && kb != kNone) { if (count_elems(kb, lb) != 0) return fail("record %lld: feature values must be float lists", rec_no); kb = kNone; }
            if (ka == kNone) la = Span{nullptr, nullptr};
            if (kb == kNone) lb = Span{nullptr, nullptr};
            const int64_t q0 = sz.nnz;
            int64_t ni = 0, nv = 0;
            Span ipacked;
            if (!o.gcol && !o.local16 && single_packed(la, ipacked)) {
                // counting pass: the indices are not decoded (their range is tracked by the filling pass)
                if (!count_varints(ipacked, ni)) return fail("record %lld: malformed index list", rec_no);
            } else {
                int64_t mn = sz.min_index, mx = sz.max_index;
                if (o.local16) {
                    if (!each_int64(la, [&](int64_t v) { ent_g_.push_back(v); mn = v < mn ? v : mn; mx = v > mx ? v : mx; ni++; }))
                        return fail("record %lld: malformed index list", rec_no);
                } else if (!each_int64(la, [&](int64_t v) { if (o.gcol) o.gcol[q0 + ni] = v; mn = v < mn ? v : mn; mx = v > mx ? v : mx; ni++; }))
                    return fail("record %lld: malformed index list", rec_no);
                sz.min_index = mn; sz.max_index = mx;
            }
            Span packed;
            if (single_packed(lb, packed) && (packed.e - packed.p) % 4 == 0) {
                nv = (packed.e - packed.p) / 4;
                if (o.val && nv) memcpy(o.val + q0, packed.p, 4 * (size_t)(nv < ni ? nv : ni));
            } else if (!each_float(lb, [&](float v) { if (o.val && nv < ni) o.val[q0 + nv] = v; nv++; }))
                return fail("record %lld: malformed value list", rec_no);
            if (ni != nv) return fail("record %lld: indices / values length mismatch (%lld vs %lld)", rec_no, ni, nv);
            if (si <= n && o.row_len) o.row_len[row0 + si - 1] = ni;
            sz.nnz += ni;
        }
        if (si != n || sv != n)
            return fail("record %lld: %lld index lists / %lld value lists for its samples", rec_no, si, sv);
        if (o.local16) {
            if (!local_index(o, ent_q0, sz.n_entities)) return fail("record %lld: more than 65535 distinct features (fused local index)", rec_no);
        }
        sz.n_rows += n;
        sz.n_entities++;
        return true;
    }

    // np.unique(cols, return_inverse=True) of the entity just parsed (job_consumers.py:243) while its indices are
    // still in cache: one open-addressing table per reader, two probes per non-zero (local_index_pass1's algorithm).
    // Negative indices hash like any other (the caller rejects the partition by the index range).
    bool local_index(const SeqexOut &o, const int64_t q0, const int64_t e)
    {
        const size_t nz = ent_g_.size();
        if (nz == 0) { o.d_e[e] = 0; return true; }
        size_t cap = 16;
        while (cap < 2 * nz) cap <<= 1;
        if (keys_.size() < cap) { keys_.assign(cap, INT64_MIN); slot_rank_.assign(cap, 0); }
        const size_t mask = keys_.size() - 1;
        auto hash = [](int64_t g) { return (size_t)((uint64_t)g * 0x9E3779B97F4A7C15ull >> 20); };
        distinct_.clear();
        slots_.clear();
        for (size_t j = 0; j < nz; j++) {
            const int64_t g = ent_g_[j];
            size_t h = hash(g) & mask;
            while (keys_[h] != INT64_MIN && keys_[h] != g) h = (h + 1) & mask;
            if (keys_[h] == INT64_MIN) { keys_[h] = g; distinct_.push_back(g); slots_.push_back((uint32_t)h); }
        }
        bool ok = distinct_.size() <= 65535;
        if (ok) {
            std::sort(distinct_.begin(), distinct_.end());
            for (size_t r = 0; r < distinct_.size(); r++) {
                const int64_t g = distinct_[r];
                size_t h = hash(g) & mask;
                while (keys_[h] != g) h = (h + 1) & mask;
                slot_rank_[h] = (int32_t)r;
                o.uniq_scratch[q0 + (int64_t)r] = g;
            }
            for (size_t j = 0; j < nz; j++) {
                const int64_t g = ent_g_[j];
                size_t h = hash(g) & mask;
                while (keys_[h] != g) h = (h + 1) & mask;
                o.local16[q0 + (int64_t)j] = (uint16_t)slot_rank_[h];
            }
            o.d_e[e] = (int64_t)distinct_.size();
        }
        for (const uint32_t h : slots_) keys_[h] = INT64_MIN;     // reset only what was touched
        return ok;
    }

    const gdmix_seqex_spec &spec_;
    std::string &err_;
    std::vector<int64_t> ent_g_, keys_, distinct_;
    std::vector<int32_t> slot_rank_;
    std::vector<uint32_t> slots_;
};

// ---------------------------------------------------------------------------------------------------------
// Fixed-effect input: one tf.train.Example per row (per_record_input_fn, input_data_pipeline.py:223-243): the bag's
// indices / values as two lists of the features map, and scalar uid / label / offset / weight columns.
// ---------------------------------------------------------------------------------------------------------
struct ExampleOut {   // null pointers: counting pass
    int64_t *row_len = nullptr, *uid = nullptr;
    int32_t *col = nullptr;
    float *val = nullptr, *label = nullptr, *offset = nullptr, *weight = nullptr;
};

class ExampleReader {
public:
    ExampleReader(const gdmix_seqex_spec &spec, std::string &err) : spec_(spec), err_(err) {}

    bool run(const uint8_t *buf, int64_t len, gdmix_seqex_sizes &sz, const ExampleOut &o)
    {
        memset(&sz, 0, sizeof(sz));
        sz.all_labelled = 1;
        const uint8_t *p = buf, *end = buf + len;
        while (p < end) {
            if (end - p < 12) return fail("truncated TFRecord header at byte %lld", (long long)(p - buf));
            uint64_t n; memcpy(&n, p, 8);
            if (n > (uint64_t)(end - p - 12) || (uint64_t)(end - p - 12) - n < 4)
                return fail("truncated TFRecord payload at byte %lld", (long long)(p - buf));
            if (!record(Span{p + 12, p + 12 + n}, sz, o)) return false;
            p += 12 + n + 4;
        }
        return true;
    }

private:
    bool fail(const char *fmt, long long a = 0, long long b = 0)
    {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), fmt, a, b);
        err_ = tmp;
        return false;
    }
    static bool name_is(Span key, const char *name)
    {
        if (!name) return false;
        const size_t n = strlen(name);
        return (size_t)(key.e - key.p) == n && memcmp(key.p, name, n) == 0;
    }
    // first element of a numeric list as double; false when the column is absent or empty
    static bool first_number(bool present, Span f, double &out, bool &malformed)
    {
        if (!present) return false;
        Kind k; Span l;
        if (!feature_kind(f, k, l)) { malformed = true; return false; }
        bool got = false;
        if (k == kFloat) { if (!each_float(l, [&](float v) { if (!got) { out = (double)v; got = true; } })) malformed = true; }
        else if (k == kInt64) { if (!each_int64(l, [&](int64_t v) { if (!got) { out = (double)v; got = true; } })) malformed = true; }
        else if (k == kBytes) malformed = true;
        return got;
    }

    bool record(Span rec, gdmix_seqex_sizes &sz, const ExampleOut &o)
    {
        const int64_t row = sz.n_rows;
        Span features{nullptr, nullptr};
        uint32_t fno, wt; uint64_t sc; Span sub;
        while (!rec.empty()) {
            if (!rd_field(rec, fno, wt, sc, sub)) return fail("row %lld: malformed Example", row);
            if (wt == 2 && fno == 1) { features = sub; break; }
        }
        Span f_idx{nullptr, nullptr}, f_val = f_idx, f_uid = f_idx, f_label = f_idx, f_off = f_idx, f_w = f_idx;
        bool h_idx = false, h_val = false, h_uid = false, h_label = false, h_off = false, h_w = false;
        Span c = features;
        while (!c.empty()) {
            if (!rd_field(c, fno, wt, sc, sub)) return fail("row %lld: malformed features map", row);
            if (fno != 1 || wt != 2) continue;
            Span key{nullptr, nullptr}, val{nullptr, nullptr};
            Span entry = sub;
            while (!entry.empty()) {
                uint32_t f2, w2; uint64_t s2; Span sub2;
                if (!rd_field(entry, f2, w2, s2, sub2)) return fail("row %lld: malformed features entry", row);
                if (w2 == 2 && f2 == 1) key = sub2;
                else if (w2 == 2 && f2 == 2) val = sub2;
            }
            if (name_is(key, spec_.bag_indices)) { f_idx = val; h_idx = true; }
            else if (name_is(key, spec_.bag_values)) { f_val = val; h_val = true; }
            else if (name_is(key, spec_.uid)) { f_uid = val; h_uid = true; }
            else if (name_is(key, spec_.label)) { f_label = val; h_label = true; }
            else if (name_is(key, spec_.offset)) { f_off = val; h_off = true; }
            else if (name_is(key, spec_.weight)) { f_w = val; h_w = true; }
        }
        int64_t ni = 0, nv = 0;
        if (spec_.bag_indices) {
            Kind ka = kNone, kb = kNone; Span la{nullptr, nullptr}, lb{nullptr, nullptr};
            if (h_idx && !feature_kind(f_idx, ka, la)) return fail("row %lld: malformed index feature", row);
            if (h_val && !feature_kind(f_val, kb, lb)) return fail("row %lld: malformed value feature", row);
            if (ka != kInt64 && ka != kNone) { if (count_elems(ka, la) != 0) return fail("row %lld: feature indices must be an int64 list", row); la = Span{nullptr, nullptr}; }
            if (kb != kFloat && kb != kNone) { if (count_elems(kb, lb) != 0) return fail("row %lld: feature values must be a float list", row); lb = Span{nullptr, nullptr}; }
            const int64_t q0 = sz.nnz;
            bool range_ok = true;
            if (!each_int64(la, [&](int64_t v) { if (v < 0 || v > 0x7fffffffLL) range_ok = false; if (o.col) o.col[q0 + ni] = (int32_t)v; ni++; }))
                return fail("row %lld: malformed index list", row);
            if (!range_ok) return fail("row %lld: feature index outside int32", row);
            if (!each_float(lb, [&](float v) { if (o.val && nv < ni) o.val[q0 + nv] = v; nv++; }))
                return fail("row %lld: malformed value list", row);
            if (ni != nv) return fail("row %lld: indices / values length mismatch (%lld indices)", row, ni);
            sz.nnz += ni;
        }
        if (o.row_len) o.row_len[row] = ni;
        bool bad = false;
        double v;
        const bool g_uid = first_number(h_uid, f_uid, v, bad);
        if (o.uid) o.uid[row] = g_uid ? (int64_t)v : 0;
        if (g_uid && o.uid) {   // uids are int64: take the exact value, not its double
            Kind k; Span l; feature_kind(f_uid, k, l);
            if (k == kInt64) { bool got = false; each_int64(l, [&](int64_t x) { if (!got) { o.uid[row] = x; got = true; } }); }
        }
        const bool g_label = first_number(h_label, f_label, v, bad);
        if (o.label) { if (g_label) o.label[row] = (float)v; else { const uint32_t nanbits = 0x7fc00000u; memcpy(&o.label[row], &nanbits, 4); } }
        if (!g_label) sz.all_labelled = 0;
        const bool g_off = first_number(h_off, f_off, v, bad);
        if (o.offset) o.offset[row] = g_off ? (float)v : 0.0f;
        const bool g_w = first_number(h_w, f_w, v, bad);
        if (o.weight) o.weight[row] = g_w ? (float)v : 1.0f;
        if (h_w) sz.saw_weight = 1;
        if (bad) return fail("row %lld: malformed scalar column", row);
        sz.n_rows++;
        sz.n_entities = sz.n_rows;
        return true;
    }

    const gdmix_seqex_spec &spec_;
    std::string &err_;
};

// ---------------------------------------------------------------------------------------------------------
// Entity-local feature indexing on the host (np.unique(cols, return_inverse=True) per entity,
// job_consumers.py:243): entity e owns rows [ent_rowptr[e], ent_rowptr[e+1]) and their non-zeros; its distinct
// global ids, ascending, become local indices 0 .. d_e - 1.  All host threads, one entity at a time each:
// an open-addressing table of the entity's ids, the distinct ones sorted, ranks written back.
//   pass 1 (uniq_global == nullptr): local[nnz], d_e[E], and the distinct ids parked in `scratch` (int64[nnz], entity
//           e's at its first non-zero's position)
//   pass 2: uniq_global[uniq_ptr[e] ..] = the parked ids (uniq_ptr = exclusive scan of d_e, the caller's)
// Returns -1 when an id is negative.
// ---------------------------------------------------------------------------------------------------------
inline int local_index_pass1(const int64_t *ent_rowptr, const int64_t *rowptr, const int64_t *gcol, int64_t E, int32_t *local,
                             int64_t *d_e, int64_t *scratch)
{
    int bad = 0;
#pragma omp parallel reduction(| : bad)
    {
        std::vector<int64_t> keys;
        std::vector<int32_t> slot_rank;
        std::vector<int64_t> distinct;
#pragma omp for schedule(dynamic, 16)
        for (int64_t e = 0; e < E; e++) {
            const int64_t q0 = rowptr[ent_rowptr[e]], q1 = rowptr[ent_rowptr[e + 1]];
            const int64_t nz = q1 - q0;
            if (nz == 0) { d_e[e] = 0; continue; }
            size_t cap = 16;
            while (cap < (size_t)(2 * nz)) cap <<= 1;
            keys.assign(cap, -1);
            slot_rank.assign(cap, 0);
            distinct.clear();
            const size_t mask = cap - 1;
            for (int64_t q = q0; q < q1; q++) {
                const int64_t g = gcol[q];
                if (g < 0) { bad |= 1; continue; }
                size_t h = (size_t)((uint64_t)g * 0x9E3779B97F4A7C15ull >> 20) & mask;
                while (keys[h] != -1 && keys[h] != g) h = (h + 1) & mask;
                if (keys[h] == -1) { keys[h] = g; distinct.push_back(g); }
            }
            std::sort(distinct.begin(), distinct.end());
            for (size_t r = 0; r < distinct.size(); r++) {
                const int64_t g = distinct[r];
                size_t h = (size_t)((uint64_t)g * 0x9E3779B97F4A7C15ull >> 20) & mask;
                while (keys[h] != g) h = (h + 1) & mask;
                slot_rank[h] = (int32_t)r;
                scratch[q0 + (int64_t)r] = g;
            }
            for (int64_t q = q0; q < q1; q++) {
                const int64_t g = gcol[q];
                if (g < 0) { local[q] = 0; continue; }
                size_t h = (size_t)((uint64_t)g * 0x9E3779B97F4A7C15ull >> 20) & mask;
                while (keys[h] != g) h = (h + 1) & mask;
                local[q] = slot_rank[h];
            }
            d_e[e] = (int64_t)distinct.size();
        }
    }
    return bad ? -1 : 0;
}

inline void local_index_pass2(const int64_t *ent_rowptr, const int64_t *rowptr, int64_t E, const int64_t *d_e,
                              const int64_t *uniq_ptr, const int64_t *scratch, int64_t *uniq_global)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t e = 0; e < E; e++) {
        const int64_t q0 = rowptr[ent_rowptr[e]];
        for (int64_t r = 0; r < d_e[e]; r++) uniq_global[uniq_ptr[e] + r] = scratch[q0 + r];
    }
}

}  // namespace gdmix_host
