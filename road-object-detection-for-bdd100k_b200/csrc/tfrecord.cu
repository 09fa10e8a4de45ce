// f-4: the ground-truth side of the on-disk format.  The reference stores one tf.train.Example per image in
// TFRecord files (writer: dataset/pascalvoc_to_tfrecords.py:128-170, schema: dataset/pascalvoc_common.py:75-98) and
// reads them through slim's DatasetDataProvider on CPU threads.  Here a native host parser extracts the box-level
// fields of every record ONCE (ragged arrays: ymin / xmin / ymax / xmax / label / difficult / truncated + offsets),
// the arrays are uploaded to HBM once (BDD100K: ~40 MB), and every batch is then assembled on the device by
// gt_gather_kernel from a list of record indices into the padded [B,G,4] / [B,G] / counts form that
// cornerBboxes_2_centerBboxes + refine_groundtruth(gt_counts=...) take: no host work per step.
//
// Formats (public specifications):
//   TFRecord framing (tensorflow/core/lib/io/record_writer.cc): uint64 length | uint32 masked_crc32c(length) |
//     data | uint32 masked_crc32c(data), little endian; masked(c) = ((c >> 15) | (c << 17)) + 0xa282ead8, CRC-32C.
//   Example (tensorflow/core/example/{example,feature}.proto): Example{Features features = 1},
//     Features{map<string, Feature> feature = 1}, Feature{oneof{BytesList = 1, FloatList = 2, Int64List = 3}},
//     FloatList{repeated float value = 1 [packed]}, Int64List{repeated int64 value = 1 [packed]}.
#include <string.h>

#include "common.cuh"

namespace rod {

// ------------------------------------------------------------------------------------------- CRC-32C
struct CrcTable {
  uint32_t t[256];
  CrcTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[i] = c;
    }
  }
};
static const uint32_t* crc_table() {
  static const CrcTable table;                    // (thread-safe initialisation)
  return table.t;
}
static uint32_t masked_crc32c(const unsigned char* p, size_t n) {
  const uint32_t* table = crc_table();
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
  c ^= 0xFFFFFFFFu;
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

// ------------------------------------------------------------------------------------------- protobuf wire format
struct Span {
  const unsigned char* p;
  const unsigned char* end;
  bool ok;
};
static uint64_t varint(Span& s) {
  uint64_t v = 0;
  for (int shift = 0; shift < 64 && s.p < s.end; shift += 7) {
    const unsigned char b = *s.p++;
    v |= (uint64_t)(b & 0x7Fu) << shift;
    if (!(b & 0x80u)) return v;
  }
  s.ok = false;
  return 0;
}
static Span sub(Span& s) {                        // length-delimited payload
  const uint64_t n = varint(s);
  Span r{s.p, s.p, s.ok};
  if (!s.ok || n > (uint64_t)(s.end - s.p)) { s.ok = false; r.ok = false; return r; }
  r.end = s.p + n;
  s.p += n;
  return r;
}
static void skip(Span& s, unsigned wire) {
  switch (wire) {
    case 0: (void)varint(s); break;
    case 1: if (s.end - s.p >= 8) s.p += 8; else s.ok = false; break;
    case 2: (void)sub(s); break;
    case 5: if (s.end - s.p >= 4) s.p += 4; else s.ok = false; break;
    default: s.ok = false;
  }
}

// the fields of the reference schema the box-level path consumes (dataset/pascalvoc_common.py:82-97)
enum Key { K_YMIN, K_XMIN, K_YMAX, K_XMAX, K_LABEL, K_DIFFICULT, K_TRUNCATED, K_SHAPE, K_OTHER };
static Key key_of(const unsigned char* p, size_t n) {
  static const struct { const char* name; Key k; } table[] = {
      {"image/object/bbox/ymin", K_YMIN},           {"image/object/bbox/xmin", K_XMIN},
      {"image/object/bbox/ymax", K_YMAX},           {"image/object/bbox/xmax", K_XMAX},
      {"image/object/bbox/label", K_LABEL},         {"image/object/bbox/difficult", K_DIFFICULT},
      {"image/object/bbox/truncated", K_TRUNCATED}, {"image/shape", K_SHAPE}};
  for (const auto& e : table)
    if (strlen(e.name) == n && memcmp(e.name, p, n) == 0) return e.k;
  return K_OTHER;
}

struct GtOut {                                    // all optional (NULL: count only)
  float* coord[4];                                // ymin, xmin, ymax, xmax
  int64_t* ints[3];                               // label, difficult, truncated
  int64_t* shape;                                 // [records][3]
  int64_t cap;                                    // capacity of the ragged arrays (objects)
};

// values of a FloatList / Int64List payload (field 1, packed or not) -> dst[0..); returns the count, -1 on error
static int64_t read_floats(Span list, float* dst, int64_t room) {
  int64_t n = 0;
  while (list.ok && list.p < list.end) {
    const uint64_t tag = varint(list);
    const unsigned field = (unsigned)(tag >> 3), wire = (unsigned)(tag & 7u);
    if (field == 1 && wire == 2) {
      Span pk = sub(list);
      if (!pk.ok || (pk.end - pk.p) % 4) return -1;
      for (; pk.p < pk.end; pk.p += 4, ++n)
        if (dst) { if (n >= room) return -1; memcpy(dst + n, pk.p, 4); }
    } else if (field == 1 && wire == 5) {
      if (list.end - list.p < 4) return -1;
      if (dst) { if (n >= room) return -1; memcpy(dst + n, list.p, 4); }
      list.p += 4; ++n;
    } else {
      skip(list, wire);
    }
  }
  return list.ok ? n : -1;
}
static int64_t read_ints(Span list, int64_t* dst, int64_t room) {
  int64_t n = 0;
  while (list.ok && list.p < list.end) {
    const uint64_t tag = varint(list);
    const unsigned field = (unsigned)(tag >> 3), wire = (unsigned)(tag & 7u);
    if (field == 1 && wire == 2) {
      Span pk = sub(list);
      while (pk.ok && pk.p < pk.end) {
        const int64_t v = (int64_t)varint(pk);
        if (dst) { if (n >= room) return -1; dst[n] = v; }
        ++n;
      }
      if (!pk.ok) return -1;
    } else if (field == 1 && wire == 0) {
      const int64_t v = (int64_t)varint(list);
      if (dst) { if (n >= room) return -1; dst[n] = v; }
      ++n;
    } else {
      skip(list, wire);
    }
  }
  return list.ok ? n : -1;
}

// one Example: appends its objects at position `at`; returns the number of objects, -1 malformed, -2 inconsistent
static int64_t parse_example(const unsigned char* p, size_t n, int64_t at, int64_t record, const GtOut& out) {
  Span ex{p, p + n, true};
  int64_t cnt[7] = {-1, -1, -1, -1, -1, -1, -1};
  while (ex.ok && ex.p < ex.end) {
    const uint64_t tag = varint(ex);
    if ((tag >> 3) != 1 || (tag & 7u) != 2) { skip(ex, (unsigned)(tag & 7u)); continue; }
    Span feats = sub(ex);                         // Features
    while (feats.ok && feats.p < feats.end) {
      const uint64_t t2 = varint(feats);
      if ((t2 >> 3) != 1 || (t2 & 7u) != 2) { skip(feats, (unsigned)(t2 & 7u)); continue; }
      Span entry = sub(feats);                    // map entry {key = 1, value = 2}
      Key key = K_OTHER;
      Span value{nullptr, nullptr, false};
      while (entry.ok && entry.p < entry.end) {
        const uint64_t t3 = varint(entry);
        const unsigned f3 = (unsigned)(t3 >> 3), w3 = (unsigned)(t3 & 7u);
        if (f3 == 1 && w3 == 2) { Span k = sub(entry); if (k.ok) key = key_of(k.p, (size_t)(k.end - k.p)); }
        else if (f3 == 2 && w3 == 2) value = sub(entry);
        else skip(entry, w3);
      }
      if (!entry.ok) return -1;
      if (key == K_OTHER || !value.ok) continue;
      while (value.ok && value.p < value.end) {   // Feature: one of the three lists
        const uint64_t t4 = varint(value);
        const unsigned f4 = (unsigned)(t4 >> 3), w4 = (unsigned)(t4 & 7u);
        if (w4 != 2) { skip(value, w4); continue; }
        Span list = sub(value);
        if (!list.ok) return -1;
        if (key <= K_XMAX && f4 == 2) {
          cnt[key] = read_floats(list, out.coord[key] ? out.coord[key] + at : nullptr, out.cap - at);
          if (cnt[key] < 0) return -1;
        } else if (key >= K_LABEL && key <= K_TRUNCATED && f4 == 3) {
          int64_t* dst = out.ints[key - K_LABEL];
          cnt[key] = read_ints(list, dst ? dst + at : nullptr, out.cap - at);
          if (cnt[key] < 0) return -1;
        } else if (key == K_SHAPE && f4 == 3) {
          int64_t tmp[3] = {0, 0, 0};
          if (read_ints(list, tmp, 3) < 0) return -1;
          if (out.shape) memcpy(out.shape + 3 * record, tmp, sizeof(tmp));
        }
      }
      if (!value.ok) return -1;
    }
    if (!feats.ok) return -1;
  }
  if (!ex.ok) return -1;
  // VarLenFeature: an absent key is an empty list; the four coordinate lists and the labels must agree
  const int64_t g = cnt[K_YMIN] < 0 ? 0 : cnt[K_YMIN];
  if (g > out.cap - at) return -1;                 // more objects than the caller's arrays hold
  for (int k = K_XMIN; k <= K_LABEL; ++k)
    if ((cnt[k] < 0 ? 0 : cnt[k]) != g) return -2;
  for (int k = K_DIFFICULT; k <= K_TRUNCATED; ++k) {
    const int64_t c = cnt[k] < 0 ? 0 : cnt[k];
    if (c != g && c != 0) return -2;
    if (c == 0 && out.ints[k - K_LABEL])           // optional lists default to 0
      for (int64_t i = 0; i < g; ++i) out.ints[k - K_LABEL][at + i] = 0;
  }
  return g;
}

static int walk_records(const void* data, size_t n_bytes, int verify_crc, const GtOut& out, int64_t max_records,
                        int64_t* offsets, int64_t* n_records, int64_t* n_objects) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  size_t pos = 0;
  int64_t rec = 0, obj = 0;
  while (pos < n_bytes) {
    ROD_REQUIRE(n_bytes - pos >= 12, "TFRecord: truncated header at byte %zu", pos);
    uint64_t len;
    uint32_t crc;
    memcpy(&len, p + pos, 8);
    memcpy(&crc, p + pos + 8, 4);
    ROD_REQUIRE(!verify_crc || masked_crc32c(p + pos, 8) == crc, "TFRecord: corrupted length of record %lld", (long long)rec);
    ROD_REQUIRE(len <= n_bytes - pos - 12 && n_bytes - pos - 12 - len >= 4, "TFRecord: truncated record %lld", (long long)rec);
    const unsigned char* body = p + pos + 12;
    memcpy(&crc, body + len, 4);
    ROD_REQUIRE(!verify_crc || masked_crc32c(body, (size_t)len) == crc, "TFRecord: corrupted data in record %lld", (long long)rec);
    ROD_REQUIRE(max_records < 0 || rec < max_records, "TFRecord: more than %lld records", (long long)max_records);
    if (offsets) offsets[rec] = obj;
    const int64_t g = parse_example(body, (size_t)len, obj, rec, out);
    ROD_REQUIRE(g != -1, "TFRecord: record %lld is not a well-formed tf.train.Example (or holds more objects than announced)", (long long)rec);
    ROD_REQUIRE(g != -2, "TFRecord: record %lld: bbox coordinate / label lists differ in length", (long long)rec);
    obj += g;
    ++rec;
    pos += 12 + (size_t)len + 4;
  }
  if (offsets) offsets[rec] = obj;
  if (n_records) *n_records = rec;
  if (n_objects) *n_objects = obj;
  return ROD_OK;
}

// ------------------------------------------------------------------------------------------- device gather
// one warp per image of the batch: record indices[b] -> rows of the padded batch, zero padded, counts clipped to gmax;
// an index outside [0, n_records) gives an all-zero row and counts[b] = -1 (the caller may check without a round trip per step)
__global__ void __launch_bounds__(256)
gt_gather_kernel(const float* __restrict__ ymin, const float* __restrict__ xmin, const float* __restrict__ ymax,
                 const float* __restrict__ xmax, const long long* __restrict__ label, const long long* __restrict__ difficult,
                 const long long* __restrict__ offsets, const long long* __restrict__ indices, long long n_records, int batch,
                 int gmax, float* __restrict__ bboxes, long long* __restrict__ labels, long long* __restrict__ difficults,
                 int32_t* __restrict__ counts) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= batch) return;
  const long long rec = indices ? indices[b] : b;
  const bool bad = rec < 0 || rec >= n_records;
  const long long o0 = bad ? 0 : offsets[rec];
  const int g = bad ? 0 : (int)min((long long)gmax, offsets[rec + 1] - o0);
  for (int i = lane; i < gmax; i += 32) {
    const bool in = i < g;
    const float4 box = in ? make_float4(ymin[o0 + i], xmin[o0 + i], ymax[o0 + i], xmax[o0 + i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    st4(bboxes + 4ll * ((long long)b * gmax + i), box);
    labels[(long long)b * gmax + i] = in ? label[o0 + i] : 0;
    if (difficults) difficults[(long long)b * gmax + i] = (in && difficult) ? difficult[o0 + i] : 0;
  }
  if (lane == 0) counts[b] = bad ? -1 : g;
}

}  // namespace rod

extern "C" int rod_tfrecord_index(const void* data, size_t n_bytes, int verify_crc, int64_t* n_records, int64_t* n_objects) {
  using namespace rod;
  ROD_REQUIRE(data != nullptr || n_bytes == 0, "rod_tfrecord_index: NULL data");
  GtOut none = {};
  none.cap = INT64_MAX;
  return walk_records(data, n_bytes, verify_crc, none, -1, nullptr, n_records, n_objects);
}

extern "C" int rod_tfrecord_read_gt(const void* data, size_t n_bytes, int verify_crc, int64_t n_records, int64_t n_objects,
                                    float* ymin, float* xmin, float* ymax, float* xmax, int64_t* label, int64_t* difficult,
                                    int64_t* truncated, int64_t* offsets, int64_t* shape) {
  using namespace rod;
  ROD_REQUIRE(data != nullptr || n_bytes == 0, "rod_tfrecord_read_gt: NULL data");
  ROD_REQUIRE(n_records >= 0 && n_objects >= 0 && offsets != nullptr, "rod_tfrecord_read_gt: bad arguments");
  ROD_REQUIRE(n_objects == 0 || (ymin && xmin && ymax && xmax && label), "rod_tfrecord_read_gt: NULL output array");
  GtOut out = {};
  out.coord[0] = ymin; out.coord[1] = xmin; out.coord[2] = ymax; out.coord[3] = xmax;
  out.ints[0] = label; out.ints[1] = difficult; out.ints[2] = truncated;
  out.shape = shape;
  out.cap = n_objects;
  int64_t nr = 0, no = 0;
  const int rc = walk_records(data, n_bytes, verify_crc, out, n_records, offsets, &nr, &no);
  if (rc) return rc;
  ROD_REQUIRE(nr == n_records && no == n_objects, "rod_tfrecord_read_gt: found %lld records / %lld objects, expected %lld / %lld",
              (long long)nr, (long long)no, (long long)n_records, (long long)n_objects);
  return ROD_OK;
}

extern "C" int rod_gt_gather(const float* ymin, const float* xmin, const float* ymax, const float* xmax, const int64_t* label,
                             const int64_t* difficult, const int64_t* offsets, const int64_t* indices, int64_t n_records,
                             int batch, int gmax, float* bboxes, int64_t* labels, int64_t* difficults, int32_t* counts, void* stream) {
  using namespace rod;
  ROD_REQUIRE(offsets && bboxes && labels && counts, "rod_gt_gather: NULL pointer argument");
  ROD_REQUIRE(batch >= 0 && gmax >= 1 && n_records >= 0, "rod_gt_gather: batch=%d gmax=%d n_records=%lld invalid", batch, gmax,
              (long long)n_records);
  ROD_REQUIRE(indices != nullptr || batch <= n_records, "rod_gt_gather: batch=%d exceeds the %lld records", batch, (long long)n_records);
  if (batch == 0) return ROD_OK;
  gt_gather_kernel<<<(batch + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
      ymin, xmin, ymax, xmax, reinterpret_cast<const long long*>(label), reinterpret_cast<const long long*>(difficult),
      reinterpret_cast<const long long*>(offsets), reinterpret_cast<const long long*>(indices), (long long)n_records, batch, gmax,
      bboxes,
      reinterpret_cast<long long*>(labels), reinterpret_cast<long long*>(difficults), counts);
  ROD_LAUNCH_CHECK("gt_gather_kernel");
  return ROD_OK;
}
