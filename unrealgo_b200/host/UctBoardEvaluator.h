// Drop-in for search/UctBoardEvaluator.{h,cpp} (reference .cpp:6-33): config -> evaluator wiring + forwarding.
#ifndef UGEVAL_HOST_UCTBOARDEVALUATOR_H
#define UGEVAL_HOST_UCTBOARDEVALUATOR_H

#include <string>
#include <vector>

#include "DlConfig.h"
#include "DlTFNetworkEvaluator.h"

typedef double UctValueType;  // search/UctValue.h:14

class UctBoardEvaluator {
 public:
  UctBoardEvaluator() : m_evaluator("", DF_HWC) {  // .cpp:6-11
    m_evaluator.SetNetworkInput(DlConfig::GetInstance().get_network_input());
    std::vector<std::string> outputs;
    DlConfig::GetInstance().get_network_outputs(outputs);
    m_evaluator.SetNetworkOutput(outputs);
  }
  explicit UctBoardEvaluator(const std::string& graphPath, const std::string& checkpoint) : m_evaluator(graphPath) {
    if (!checkpoint.empty()) m_evaluator.UpdateCheckPoint(checkpoint);  // .cpp:13-16
  }
  void LoadGraph(const std::string& graphPath) { m_evaluator.LoadGraph(graphPath); }
  void UpdateCheckPoint(const std::string& checkpoint) { m_evaluator.UpdateCheckPoint(checkpoint); }
  void EvaluateState(char feature[][NUM_MAPS][GO_MAX_SIZE][GO_MAX_SIZE], UctValueType actions_[][GO_MAX_MOVES],
                     UctValueType value_[], int numBatches = MAX_BATCHES) {
    m_evaluator.Evaluate(feature, actions_, value_, numBatches);
  }
  bool GraphLoaded() { return m_evaluator.MetaGraphLoaded(); }
  tensorflow::DlTFNetworkEvaluator<bool>& Evaluator() { return m_evaluator; }

 private:
  tensorflow::DlTFNetworkEvaluator<bool> m_evaluator;
};
#endif
