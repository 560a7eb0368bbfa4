// Drop-in translation unit for funcapproximator/DlTFRecordWriter.cc: defines the members that the reference's
// UNMODIFIED funcapproximator/DlTFRecordWriter.h declares (that header only forward-declares TensorFlow types, so it
// needs no TensorFlow to compile) on top of ug_tfrecord_* (include/ugeval.h; csrc/tfrecord.cpp writes the TFRecord
// framing and the tf.train.Example bytes itself).  Build it in place of the reference's .cc with the reference root
// on the include path; call sites — UctDeepPlayer::WriteTFRecord, search/UctDeepPlayer.cpp:362-388 — are untouched.
//
//   reference member (funcapproximator/DlTFRecordWriter.cc)        C export
//   DlTFRecordWriter(filename)                       :18-22       ug_tfrecord_open
//   ~DlTFRecordWriter (flush, close)                 :24-32       ug_tfrecord_close
//   Flush                                            :34-37       ug_tfrecord_flush
//   WriteExample(feature17, f_len, policy, n, r)     :59-76       ug_tfrecord_write_example
//   WriteExample(names...)                           :78-100      ug_tfrecord_write_example_named
// The handle is kept in the class's one data member (a pointer to an incomplete type in the header).  Like the
// reference, a writer whose file could not be opened is not reported at construction; here its writes return -1
// instead of dereferencing a null writer.  AddByteList / AddFloatList are protected helpers of the TensorFlow
// implementation with no caller outside it and are not defined.
#include <string>

#include "funcapproximator/DlTFRecordWriter.h"

#include "../../include/ugeval.h"

namespace tensorflow {
namespace io {

namespace {
inline ug_tfrecord* handle(PyRecordWriter* p) { return reinterpret_cast<ug_tfrecord*>(p); }
}

DlTFRecordWriter::DlTFRecordWriter(const string& filename)
    : m_recordWriter(reinterpret_cast<PyRecordWriter*>(ug_tfrecord_open(filename.c_str()))) {}

DlTFRecordWriter::~DlTFRecordWriter() {
  if (m_recordWriter) ug_tfrecord_close(handle(m_recordWriter));
}

void DlTFRecordWriter::Flush() {
  if (m_recordWriter) ug_tfrecord_flush(handle(m_recordWriter));
}

int DlTFRecordWriter::WriteExample(char* feature17, size_t f_len, float* policy, size_t p_size, float reward) {
  if (!m_recordWriter) return -1;
  return ug_tfrecord_write_example(handle(m_recordWriter), feature17, f_len, policy, p_size, reward);
}

int DlTFRecordWriter::WriteExample(const string& feature_name, char* feature17, size_t f_len,
                                   const string& policy_name, float* policy, size_t p_size,
                                   const string& reward_name, float reward) {
  if (!m_recordWriter) return -1;
  return ug_tfrecord_write_example_named(handle(m_recordWriter), feature_name.c_str(), feature17, f_len,
                                         policy_name.c_str(), policy, p_size, reward_name.c_str(), reward);
}

}  // namespace io
}  // namespace tensorflow
