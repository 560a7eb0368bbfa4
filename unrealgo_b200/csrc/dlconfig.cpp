// dlcfg-<size> key=value config (DlConfig, funcapproximator/DlConfig.cc:20-41,164-246).  Behaviours kept on
// purpose: only ' ' is trimmed (lib/StringUtil.cc:48-55); a line is a comment iff its first character after
// trimming is '#'; the split is at the FIRST '='; a line without '=' maps the whole line to itself; the first
// occurrence of a key wins (std::map::insert); a missing file is an empty map; get(key, default) returns the
// default when the stored value is empty.
#include <sys/stat.h>

#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../../include/ugeval.h"

struct ug_dlcfg {
  std::map<std::string, std::string> kv;
};

namespace {
void trim(std::string& s) {
  if (s.empty()) return;
  s.erase(0, s.find_first_not_of(' '));
  s.erase(s.find_last_not_of(' ') + 1);
}
std::string get(const ug_dlcfg* c, const std::string& key, const std::string& dflt) {
  auto it = c->kv.find(key);
  std::string v = it == c->kv.end() ? std::string() : it->second;
  return v.empty() ? dflt : v;
}
}  // namespace

extern "C" {

ug_dlcfg* ug_dlcfg_open(const char* path) {
  ug_dlcfg* c = new ug_dlcfg();
  struct stat st;
  if (!path || stat(path, &st) != 0) return c;
  std::ifstream file(path, std::ios::in);
  std::string line;
  while (std::getline(file, line)) {
    trim(line);
    if (line.empty() || line[0] != '#') {
      const std::string::size_type pos = line.find('=');
      std::string key = line.substr(0, pos);
      std::string value = line.substr(pos + 1);  // npos + 1 == 0: the whole line, as in the reference
      trim(key);
      trim(value);
      c->kv.insert(std::make_pair(key, value));
    }
  }
  return c;
}

void ug_dlcfg_close(ug_dlcfg* c) { delete c; }

int ug_dlcfg_get(const ug_dlcfg* c, const char* key, const char* dflt, char* out, int cap) {
  if (!c || !key || !out || cap <= 0) return -1;
  const std::string v = get(c, key, dflt ? dflt : "");
  const int n = (int)v.size() < cap - 1 ? (int)v.size() : cap - 1;
  memcpy(out, v.data(), (size_t)n);
  out[n] = 0;
  return n;
}

int ug_dlcfg_value_transform(const ug_dlcfg* c) {
  if (!c) return UG_TR_IDENTITY;
  const std::string v = get(c, "value_transform", "0");
  int type = 0;
  try {
    type = std::stoi(v);  // the reference lets std::stoi throw on garbage; the C ABI maps that to identity
  } catch (...) {
    return UG_TR_IDENTITY;
  }
  switch (type) {
    case 1: return UG_TR_FLIP_SIGN;
    case 2: return UG_TR_FLIP_PROB;
    default: return UG_TR_IDENTITY;
  }
}

int ug_dlcfg_reuse_search_tree(const ug_dlcfg* c) {
  if (!c) return 0;
  return get(c, "reuse_searchtree", "false") == "false" ? 0 : 1;
}

int ug_dlcfg_network_outputs(const ug_dlcfg* c, char* out, int cap, int max) {
  if (!c || !out || cap <= 0 || max <= 0) return -1;
  std::string values = get(c, "nn_outputs", "");
  if (values.empty()) return 0;
  std::vector<std::string> parts;
  size_t pos;
  while ((pos = values.find(':')) != std::string::npos) {
    parts.push_back(values.substr(0, pos));
    values.erase(0, pos + 1);
  }
  parts.push_back(values);
  int n = 0;
  for (const std::string& p : parts) {
    if (n >= max) break;
    const int len = (int)p.size() < cap - 1 ? (int)p.size() : cap - 1;
    memcpy(out + (size_t)n * cap, p.data(), (size_t)len);
    out[(size_t)n * cap + len] = 0;
    ++n;
  }
  return n;
}

}  // extern "C"
