// extern "C" wrappers over the reference's OWN classes, compiled from /root/reference where they lie
// (funcapproximator/DlConfig.cc, lib/StringUtil.cc, lib/ArrayUtil.h).  Output: oracle/_ref/libugref.so.
// Test infrastructure only — used to pin the oracle restatement and the product's dlcfg parser.
#include <sys/stat.h>

#include <cstring>
#include <string>
#include <vector>

#include "funcapproximator/DlConfig.h"
#include "lib/ArrayUtil.h"
#include "lib/FileUtil.h"
#include "lib/StringUtil.h"

// the one FileUtil function DlConfig.cc needs (lib/FileUtil.cc:173-176 is a stat() wrapper; FileUtil.cc itself
// needs real Boost.Filesystem and is not compiled)
bool UnrealGo::FileExists(const std::string& fileName) {
  struct stat buffer;
  return stat(fileName.c_str(), &buffer) == 0;
}

#if __cplusplus >= 201703L
// lib/FileUtil.cc:178-184 (needed by platform/SgPlatform.cpp when the whole engine is linked, oracle/_ref/ref_search)
std::string UnrealGo::GetNativeFileName(const boost::filesystem::path& file) {
  boost::filesystem::path normalized = file;
  normalized = normalized.lexically_normal();
  return normalized.string();
}
#endif

#ifndef UGREF_FILE_EXISTS_ONLY
extern "C" {

void* ref_dlcfg_open(const char* path) { return new DlConfig(path); }
void ref_dlcfg_close(void* c) { delete static_cast<DlConfig*>(c); }
static int copy_out(const std::string& v, char* out, int cap) {
  const int n = (int)v.size() < cap - 1 ? (int)v.size() : cap - 1;
  memcpy(out, v.data(), (size_t)n);
  out[n] = 0;
  return n;
}
int ref_dlcfg_get(void* c, const char* key, const char* dflt, char* out, int cap) {
  return copy_out(static_cast<DlConfig*>(c)->get(key, dflt), out, cap);
}
int ref_dlcfg_network_input(void* c, char* out, int cap) {
  return copy_out(static_cast<DlConfig*>(c)->get_network_input(), out, cap);
}
int ref_dlcfg_metagraph(void* c, char* out, int cap) {
  return copy_out(static_cast<DlConfig*>(c)->get_metagraph(), out, cap);
}
int ref_dlcfg_checkpoint(void* c, char* out, int cap) {
  return copy_out(static_cast<DlConfig*>(c)->default_checkpoint(), out, cap);
}
int ref_dlcfg_network_outputs(void* c, char* out, int cap, int max) {
  std::vector<std::string> outs;
  static_cast<DlConfig*>(c)->get_network_outputs(outs);
  int n = 0;
  for (const std::string& s : outs) {
    if (n >= max) break;
    copy_out(s, out + (size_t)n * cap, cap);
    ++n;
  }
  return n;
}
int ref_dlcfg_value_transform(void* c) { return (int)static_cast<DlConfig*>(c)->get_value_transform(); }
int ref_dlcfg_reuse_search_tree(void* c) { return static_cast<DlConfig*>(c)->reuse_search_tree() ? 1 : 0; }

int ref_get_offset(const int* dims, int ndims, const int* idx) {
  return UnrealGo::ArrayUtil::GetOffset(std::vector<int>(dims, dims + ndims), ndims,
                                        std::vector<int>(idx, idx + ndims));
}

}  // extern "C"
#endif  // UGREF_FILE_EXISTS_ONLY
