// Drop-in for the hot-path subset of funcapproximator/DlConfig.{h,cc}: the process-wide "dlcfg-<size>" singleton
// (DlConfig.cc:13-18) and the getters the evaluator path reads (:77-79, :107-109, :210-246), over ug_dlcfg_*.
#ifndef UGEVAL_HOST_DLCONFIG_H
#define UGEVAL_HOST_DLCONFIG_H

#include <string>
#include <vector>

#include "../../include/ugeval.h"

enum ValueTransformType { TR_IDENTITY, TR_FLIP_SIGN, TR_FLIP_PROB, TR_UNKNOWN };  // DlConfig.h:8-13

class DlConfig {
 public:
  static DlConfig& GetInstance() {
    static DlConfig* instance = new DlConfig("dlcfg-" + std::to_string(UG_BOARD_SIZE));
    return *instance;
  }
  explicit DlConfig(const std::string& filename) : m_c(ug_dlcfg_open(filename.c_str())), valueTrans(TR_UNKNOWN) {}
  ~DlConfig() { ug_dlcfg_close(m_c); }
  DlConfig(const DlConfig&) = delete;
  DlConfig& operator=(const DlConfig&) = delete;

  std::string get(const std::string& key, const std::string& defaultValue = "") {
    char buf[4096];
    ug_dlcfg_get(m_c, key.c_str(), defaultValue.c_str(), buf, (int)sizeof(buf));
    return buf;
  }
  void setMetagraph(const std::string& metagraph) { current_metagraph = metagraph; }
  std::string get_metagraph() {
    if (current_metagraph.empty()) current_metagraph = get("meta_graph", "bootstrap-ckpt.meta");
    return current_metagraph;
  }
  std::string default_meta_graph() { return get("meta_graph", "bootstrap-ckpt.meta"); }
  std::string default_checkpoint() { return get("checkpoint", "bootstrap-ckpt"); }
  std::string get_network_input() { return get("nn_input", "input"); }
  void get_network_outputs(std::vector<std::string>& outputs) {
    char buf[8][512];
    const int n = ug_dlcfg_network_outputs(m_c, &buf[0][0], 512, 8);
    for (int i = 0; i < n; ++i) outputs.emplace_back(buf[i]);
  }
  bool reuse_search_tree() { return ug_dlcfg_reuse_search_tree(m_c) != 0; }
  ValueTransformType get_value_transform() {
    if (valueTrans == TR_UNKNOWN) valueTrans = (ValueTransformType)ug_dlcfg_value_transform(m_c);
    return valueTrans;
  }

 private:
  ug_dlcfg* m_c;
  std::string current_metagraph;
  ValueTransformType valueTrans;
};
#endif
