// Uses tensorflow::io::DlTFRecordWriter exactly as declared by the REFERENCE's own header
// (funcapproximator/DlTFRecordWriter.h, found on the include path under /root/reference), with the members defined
// by unrealgo_b200/host/DlTFRecordWriter.cc instead of the reference's TensorFlow-based .cc.
// Writes `n` deterministic examples to argv[1]: even records through the 5-argument WriteExample, odd ones through
// the named overload.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "funcapproximator/DlTFRecordWriter.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const int n = atoi(argv[2]);
  {
    tensorflow::io::DlTFRecordWriter writer(argv[1]);
    char feature[6137];
    float policy[362];
    unsigned s = 12345u;
    for (int r = 0; r < n; ++r) {
      for (int i = 0; i < 6137; ++i) { s = s * 1664525u + 1013904223u; feature[i] = (char)((s >> 24) & 1); }
      for (int i = 0; i < 362; ++i) { s = s * 1664525u + 1013904223u; policy[i] = (float)(s >> 8) / 16777216.0f; }
      const float reward = (r % 3) - 1.0f;
      int rc;
      if (r % 2 == 0) rc = writer.WriteExample(feature, sizeof(feature), policy, 362, reward);
      else rc = writer.WriteExample("f", feature, sizeof(feature), "p", policy, 362, "r", reward);
      if (rc != 0) return 3;
      if (r == 0) writer.Flush();
    }
  }  // destructor flushes and closes
  tensorflow::io::DlTFRecordWriter bad("/nonexistent-dir/x.tfrecord");
  float p = 0;
  char f = 0;
  printf("bad_write %d\n", bad.WriteExample(&f, 1, &p, 1, 0.f));
  return 0;
}
