// Self-play record writer: the on-disk format on the far side of the hot path.
//
// Replaces tensorflow::io::DlTFRecordWriter (funcapproximator/DlTFRecordWriter.cc:18-100), which hands a
// tf.train.Example to TensorFlow's PyRecordWriter.  TensorFlow is not a dependency here, so both layers are
// written out directly:
//   record framing (TensorFlow's RecordWriter, un-vendored): u64 length | u32 masked_crc32c(length) | payload |
//     u32 masked_crc32c(payload), little-endian, masked(c) = ((c >> 15) | (c << 17)) + 0xa282ead8;
//   payload: Example{features=1}, Features{map<string,Feature> feature=1}, Feature{bytes_list=1 | float_list=2},
//     BytesList{value=1}, FloatList{value=1, packed} — proto3 wire format.
// One example per position, keys as WriteExample uses them (.cc:59-76): "feature" = the 6137 plane bytes,
// "policy" = 362 floats, "reward" = 1 float.  Map entries are written in key order (protobuf's deterministic
// serialisation; the reference's order is the hash order of its map and not part of the format).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ugeval.h"
#include "tfmodel.h"  // crc32c

namespace {

void put_varint(std::string* s, uint64_t v) {
  while (v >= 0x80) { s->push_back((char)(v | 0x80)); v >>= 7; }
  s->push_back((char)v);
}
void put_len_field(std::string* s, int field, const std::string& payload) {
  put_varint(s, ((uint64_t)field << 3) | 2);
  put_varint(s, payload.size());
  s->append(payload);
}
std::string map_entry(const std::string& key, const std::string& feature) {
  std::string e;
  put_len_field(&e, 1, key);
  put_len_field(&e, 2, feature);
  return e;
}
std::string bytes_feature(const void* p, size_t n) {
  std::string list, f;
  put_len_field(&list, 1, std::string(static_cast<const char*>(p), n));  // BytesList.value
  put_len_field(&f, 1, list);                                            // Feature.bytes_list
  return f;
}
std::string float_feature(const float* p, size_t n) {
  std::string packed(reinterpret_cast<const char*>(p), n * sizeof(float)), list, f;  // little-endian host
  put_len_field(&list, 1, packed);  // FloatList.value, packed
  put_len_field(&f, 2, list);       // Feature.float_list
  return f;
}
uint32_t masked_crc(const void* p, size_t n) {
  const uint32_t c = ug::crc32c(static_cast<const uint8_t*>(p), n);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

}  // namespace

struct ug_tfrecord {
  FILE* f = nullptr;
  uint64_t records = 0;
};

extern "C" {

ug_tfrecord* ug_tfrecord_open(const char* path) {
  if (!path) return nullptr;
  FILE* f = fopen(path, "wb");
  if (!f) return nullptr;
  ug_tfrecord* w = new ug_tfrecord();
  w->f = f;
  return w;
}

int ug_tfrecord_write_example_named(ug_tfrecord* w, const char* feature_name, const void* feature, size_t f_len,
                                    const char* policy_name, const float* policy, size_t p_size,
                                    const char* reward_name, float reward) {
  if (!w || !w->f || !feature || !policy || !feature_name || !policy_name || !reward_name) return -1;
  struct KV { std::string k, v; };
  std::vector<KV> kv = {{feature_name, bytes_feature(feature, f_len)},
                        {policy_name, float_feature(policy, p_size)},
                        {reward_name, float_feature(&reward, 1)}};
  for (size_t i = 0; i < kv.size(); ++i)  // key order
    for (size_t j = i + 1; j < kv.size(); ++j)
      if (kv[j].k < kv[i].k) std::swap(kv[i], kv[j]);
  std::string features, example;
  for (const KV& e : kv) put_len_field(&features, 1, map_entry(e.k, e.v));
  put_len_field(&example, 1, features);
  const uint64_t len = example.size();
  uint8_t head[12];
  memcpy(head, &len, 8);
  const uint32_t lc = masked_crc(head, 8);
  memcpy(head + 8, &lc, 4);
  const uint32_t dc = masked_crc(example.data(), example.size());
  if (fwrite(head, 1, 12, w->f) != 12 || fwrite(example.data(), 1, example.size(), w->f) != example.size() ||
      fwrite(&dc, 1, 4, w->f) != 4)
    return -1;
  ++w->records;
  return 0;
}

int ug_tfrecord_write_example(ug_tfrecord* w, const void* feature17, size_t f_len, const float* policy,
                              size_t p_size, float reward) {
  return ug_tfrecord_write_example_named(w, "feature", feature17, f_len, "policy", policy, p_size, "reward", reward);
}

int ug_tfrecord_flush(ug_tfrecord* w) { return (w && w->f && fflush(w->f) == 0) ? 0 : -1; }

int64_t ug_tfrecord_close(ug_tfrecord* w) {
  if (!w) return -1;
  int64_t n = (int64_t)w->records;
  if (w->f && fclose(w->f) != 0) n = -1;
  delete w;
  return n;
}

}  // extern "C"
