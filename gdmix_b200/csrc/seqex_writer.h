// seqex_writer.h -- host-side writer of entity-grouped TFRecord files (no TensorFlow): the files DataPartitioner's Spark
// job saves for the random-effect trainer (gdmix-data/.../data/DataPartitioner.scala:203-280 -> IoUtils.saveDataFrame,
// utils/IoUtils.scala:131-156, spark-tfrecord recordType SequenceExample) and that per_entity_grouped_input_fn reads
// (gdmix-trainer/src/gdmix/io/input_data_pipeline.py:244-273).  One record = one tf.train.SequenceExample = one
// (entity, group): context holds the entity id (one int64 or one bytes value) and one list per sample column (uid
// int64, label int64 or float, offset float, weight float); feature_lists hold <bag>_indices (an int64 list per
// sample) and <bag>_values (a float list per sample).  TFRecord framing: u64 length | masked crc32c(length) | payload |
// masked crc32c(payload).  The bytes equal what gdmix_b200/io/tfrecord.py's encoder writes for the same columns in the
// same order (tests/test_native_reader.py); records are sized and written by all host threads.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace gdmix_host {

struct Crc32cTable {
    uint32_t t[8][256];
    Crc32cTable()
    {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; i++)
            for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
    }
};

inline uint32_t crc32c(const uint8_t *p, size_t n)
{
    static const Crc32cTable T;
    uint32_t c = 0xFFFFFFFFu;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = T.t[7][w & 0xff] ^ T.t[6][(w >> 8) & 0xff] ^ T.t[5][(w >> 16) & 0xff] ^ T.t[4][(w >> 24) & 0xff] ^
            T.t[3][(w >> 32) & 0xff] ^ T.t[2][(w >> 40) & 0xff] ^ T.t[1][(w >> 48) & 0xff] ^ T.t[0][(w >> 56) & 0xff];
        p += 8; n -= 8;
    }
    while (n--) c = T.t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
inline uint32_t masked_crc32c(const uint8_t *p, size_t n)
{
    const uint32_t c = crc32c(p, n);
    return ((c >> 15) | (c << 17)) + 0xA282EAD8u;
}

inline int64_t vsize(uint64_t v)
{
    int64_t n = 1;
    while (v >= 0x80) { v >>= 7; n++; }
    return n;
}
inline uint8_t *vput(uint8_t *o, uint64_t v)
{
    while (v >= 0x80) { *o++ = (uint8_t)(v | 0x80); v >>= 7; }
    *o++ = (uint8_t)v;
    return o;
}
// length-delimited field header
inline int64_t ld_size(int64_t payload) { return 1 + vsize((uint64_t)payload) + payload; }
inline uint8_t *ld_put(uint8_t *o, int field, int64_t payload)
{
    *o++ = (uint8_t)((field << 3) | 2);
    return vput(o, (uint64_t)payload);
}

struct SeqexColumns {
    const char *entity, *uid, *label, *offset, *weight, *bag_indices, *bag_values;   // names; NULL = column not written
    int64_t n_entities;
    const int64_t *ent_rows;     // [E] samples per record
    const int64_t *entity_int;   // [E] integer entity ids, or NULL when ids are strings
    const char *id_chars;        // string ids (utf-8) ...
    const int64_t *id_ptr;       // ... [E + 1]
    const int64_t *row_len;      // [N] non-zeros per sample
    const int64_t *gcol;         // [nnz] global feature ids
    const float *val;            // [nnz]
    const int64_t *uid_v;        // [N]
    const float *label_v;        // [N] or NULL
    int32_t label_as_int;        // write the label as an int64 list (the reference's `response`) instead of a float list
    const float *offset_v;       // [N] or NULL
    const float *weight_v;       // [N] or NULL
};

inline int64_t int64_feature_size(const int64_t *v, int64_t k)
{
    int64_t inner = 0;
    for (int64_t i = 0; i < k; i++) inner += vsize((uint64_t)v[i]);
    return ld_size(ld_size(inner));
}
inline int64_t intlabel_feature_size(const float *v, int64_t k)
{
    int64_t inner = 0;
    for (int64_t i = 0; i < k; i++) inner += vsize((uint64_t)(int64_t)v[i]);
    return ld_size(ld_size(inner));
}
inline int64_t float_feature_size(int64_t k) { return ld_size(ld_size(4 * k)); }
inline int64_t map_entry_size(const char *name, int64_t value_size)
{
    const int64_t body = ld_size((int64_t)strlen(name)) + ld_size(value_size);
    return ld_size(body);
}
inline uint8_t *map_entry_head(uint8_t *o, const char *name, int64_t value_size)
{
    const int64_t nl = (int64_t)strlen(name);
    o = ld_put(o, 1, ld_size(nl) + ld_size(value_size));
    o = ld_put(o, 1, nl);
    memcpy(o, name, (size_t)nl); o += nl;
    return ld_put(o, 2, value_size);
}
inline uint8_t *put_int64_feature(uint8_t *o, const int64_t *v, int64_t k)
{
    int64_t inner = 0;
    for (int64_t i = 0; i < k; i++) inner += vsize((uint64_t)v[i]);
    o = ld_put(o, 3, ld_size(inner));
    o = ld_put(o, 1, inner);
    for (int64_t i = 0; i < k; i++) o = vput(o, (uint64_t)v[i]);
    return o;
}
inline uint8_t *put_intlabel_feature(uint8_t *o, const float *v, int64_t k)
{
    int64_t inner = 0;
    for (int64_t i = 0; i < k; i++) inner += vsize((uint64_t)(int64_t)v[i]);
    o = ld_put(o, 3, ld_size(inner));
    o = ld_put(o, 1, inner);
    for (int64_t i = 0; i < k; i++) o = vput(o, (uint64_t)(int64_t)v[i]);
    return o;
}
inline uint8_t *put_float_feature(uint8_t *o, const float *v, int64_t k)
{
    o = ld_put(o, 2, ld_size(4 * k));
    o = ld_put(o, 1, 4 * k);
    memcpy(o, v, (size_t)(4 * k));
    return o + 4 * k;
}

struct SeqexSizes { int64_t context, lists, idx_list, val_list, payload; };

inline SeqexSizes seqex_sizes(const SeqexColumns &c, int64_t e, int64_t r0, int64_t q0)
{
    SeqexSizes s{0, 0, 0, 0, 0};
    const int64_t n = c.ent_rows[e];
    if (c.entity) {
        int64_t f;
        if (c.entity_int) f = int64_feature_size(c.entity_int + e, 1);
        else { const int64_t L = c.id_ptr[e + 1] - c.id_ptr[e]; f = ld_size(ld_size(L)); }
        s.context += map_entry_size(c.entity, f);
    }
    if (c.uid) s.context += map_entry_size(c.uid, int64_feature_size(c.uid_v + r0, n));
    if (c.label && c.label_v)
        s.context += map_entry_size(c.label, c.label_as_int ? intlabel_feature_size(c.label_v + r0, n) : float_feature_size(n));
    if (c.offset && c.offset_v) s.context += map_entry_size(c.offset, float_feature_size(n));
    if (c.weight && c.weight_v) s.context += map_entry_size(c.weight, float_feature_size(n));
    if (c.bag_indices) {
        int64_t q = q0;
        for (int64_t i = 0; i < n; i++) {
            const int64_t k = c.row_len[r0 + i];
            s.idx_list += ld_size(int64_feature_size(c.gcol + q, k));
            s.val_list += ld_size(float_feature_size(k));
            q += k;
        }
        s.lists += map_entry_size(c.bag_indices, s.idx_list);
        if (c.bag_values) s.lists += map_entry_size(c.bag_values, s.val_list);
    }
    s.payload = ld_size(s.context) + ld_size(s.lists);
    return s;
}

inline uint8_t *seqex_write(const SeqexColumns &c, int64_t e, int64_t r0, int64_t q0, const SeqexSizes &s, uint8_t *o)
{
    uint8_t *const start = o;
    const uint64_t len = (uint64_t)s.payload;
    memcpy(o, &len, 8);
    const uint32_t hc = masked_crc32c(o, 8);
    memcpy(o + 8, &hc, 4);
    o += 12;
    uint8_t *const payload = o;
    const int64_t n = c.ent_rows[e];
    o = ld_put(o, 1, s.context);
    if (c.entity) {
        if (c.entity_int) {
            o = map_entry_head(o, c.entity, int64_feature_size(c.entity_int + e, 1));
            o = put_int64_feature(o, c.entity_int + e, 1);
        } else {
            const int64_t L = c.id_ptr[e + 1] - c.id_ptr[e];
            o = map_entry_head(o, c.entity, ld_size(ld_size(L)));
            o = ld_put(o, 1, ld_size(L));
            o = ld_put(o, 1, L);
            memcpy(o, c.id_chars + c.id_ptr[e], (size_t)L); o += L;
        }
    }
    if (c.uid) {
        o = map_entry_head(o, c.uid, int64_feature_size(c.uid_v + r0, n));
        o = put_int64_feature(o, c.uid_v + r0, n);
    }
    if (c.label && c.label_v) {
        if (c.label_as_int) {
            o = map_entry_head(o, c.label, intlabel_feature_size(c.label_v + r0, n));
            o = put_intlabel_feature(o, c.label_v + r0, n);
        } else {
            o = map_entry_head(o, c.label, float_feature_size(n));
            o = put_float_feature(o, c.label_v + r0, n);
        }
    }
    if (c.offset && c.offset_v) { o = map_entry_head(o, c.offset, float_feature_size(n)); o = put_float_feature(o, c.offset_v + r0, n); }
    if (c.weight && c.weight_v) { o = map_entry_head(o, c.weight, float_feature_size(n)); o = put_float_feature(o, c.weight_v + r0, n); }
    o = ld_put(o, 2, s.lists);
    if (c.bag_indices) {
        o = map_entry_head(o, c.bag_indices, s.idx_list);
        int64_t q = q0;
        for (int64_t i = 0; i < n; i++) {
            const int64_t k = c.row_len[r0 + i];
            o = ld_put(o, 1, int64_feature_size(c.gcol + q, k));
            o = put_int64_feature(o, c.gcol + q, k);
            q += k;
        }
        if (c.bag_values) {
            o = map_entry_head(o, c.bag_values, s.val_list);
            q = q0;
            for (int64_t i = 0; i < n; i++) {
                const int64_t k = c.row_len[r0 + i];
                o = ld_put(o, 1, float_feature_size(k));
                o = put_float_feature(o, c.val + q, k);
                q += k;
            }
        }
    }
    const uint32_t pc = masked_crc32c(payload, (size_t)(o - payload));
    memcpy(o, &pc, 4);
    o += 4;
    (void)start;
    return o;
}

// out == NULL: *written = bytes needed.  Returns 0, or -1 when capacity is too small.
inline int seqex_encode(const SeqexColumns &c, uint8_t *out, int64_t capacity, int64_t *written)
{
    const int64_t E = c.n_entities;
    std::vector<int64_t> r0((size_t)E + 1, 0), q0((size_t)E + 1, 0), off((size_t)E + 1, 0);
    for (int64_t e = 0; e < E; e++) r0[(size_t)e + 1] = r0[(size_t)e] + c.ent_rows[e];
    {
        // non-zero offsets of the entities: prefix over row lengths (sequential: one pass over N int64)
        int64_t q = 0, r = 0;
        for (int64_t e = 0; e < E; e++) {
            q0[(size_t)e] = q;
            if (c.row_len) for (int64_t i = 0; i < c.ent_rows[e]; i++) q += c.row_len[r + i];
            r += c.ent_rows[e];
        }
        q0[(size_t)E] = q;
    }
    std::vector<SeqexSizes> sz((size_t)E);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t e = 0; e < E; e++) sz[(size_t)e] = seqex_sizes(c, e, r0[(size_t)e], q0[(size_t)e]);
    for (int64_t e = 0; e < E; e++) off[(size_t)e + 1] = off[(size_t)e] + 16 + sz[(size_t)e].payload;
    *written = off[(size_t)E];
    if (!out) return 0;
    if (capacity < off[(size_t)E]) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t e = 0; e < E; e++) seqex_write(c, e, r0[(size_t)e], q0[(size_t)e], sz[(size_t)e], out + off[(size_t)e]);
    return 0;
}

}  // namespace gdmix_host
