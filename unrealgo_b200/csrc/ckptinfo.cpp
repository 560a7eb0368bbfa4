// Checkpoint info files — the hand-off between the trainer and the self-play evaluators.
//
// DlCheckPoint::GetCheckPointInfo / WriteCheckpointInfo (funcapproximator/DlCheckPoint.cc:162-172, :188-203):
// a text file of three lines — checkpoint name (TF bundle prefix), sha1 of its zip, download url; the writer puts
// no newline after the url.  DlCheckPoint::CheckDefaultCheckPoint (:130-137): a checkpoint is usable when
// <name>.index and <name>.data-00000-of-00001 both exist.  The download side (curl, zip, minio) is out of scope;
// a caller that sees the name change hands it to ug_queue_swap_checkpoint / ug_eval_load_checkpoint, which is what
// NetworkEvalThread does with new_checkpoint (search/UctSearch.cpp:178-182).
#include <sys/stat.h>

#include <cstring>
#include <fstream>
#include <string>

#include "../../include/ugeval.h"

namespace {
bool exists(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}
void copy_out(const std::string& v, char* out, int cap) {
  if (!out || cap <= 0) return;
  const size_t n = v.size() < (size_t)cap - 1 ? v.size() : (size_t)cap - 1;
  memcpy(out, v.data(), n);
  out[n] = 0;
}
}  // namespace

extern "C" {

int ug_ckptinfo_read(const char* path, char* name, char* sha1, char* url, int cap) {
  if (!path) return -1;
  std::ifstream f(path);
  if (!f) return 1;  // GetCheckPointInfo leaves the struct untouched when the file is missing
  std::string l[3];
  int n = 0;
  // UnrealGo::ReadLines: one entry per line; fewer than three lines => nothing is taken (:166)
  while (n < 3 && std::getline(f, l[n])) {
    if (!l[n].empty() && l[n].back() == '\r') l[n].pop_back();
    ++n;
  }
  if (n < 3) return 1;
  copy_out(l[0], name, cap);
  copy_out(l[1], sha1, cap);
  copy_out(l[2], url, cap);
  return 0;
}

int ug_ckptinfo_write(const char* path, const char* name, const char* sha1, const char* url, int override_if_exists) {
  if (!path || !name) return -1;
  if (exists(path) && !override_if_exists) return 1;  // :196-197
  std::ofstream f(path, std::ios::out);
  if (!f) return -1;
  f << name << "\n" << (sha1 ? sha1 : "") << "\n" << (url ? url : "");
  return f.good() ? 0 : -1;
}

int ug_checkpoint_files_exist(const char* prefix) {
  if (!prefix || !prefix[0]) return 0;
  const std::string p(prefix);
  return exists(p + ".index") && exists(p + ".data-00000-of-00001") ? 1 : 0;
}

}  // extern "C"
