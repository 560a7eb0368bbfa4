// Include-path overlay for building the reference tree against ugeval: put `-I <repo>/unrealgo_b200/host/overlay`
// AHEAD of the reference root on the compiler command line and every `#include "funcapproximator/DlTFNetworkEvaluator.h"`
// in the reference (search/UctBoardEvaluator.h:12, search/UctBoardEvaluator.cpp:3) resolves to this file instead of
// the TensorFlow-based header.  The reference's own BoardStaticConfig.h, DataFormat.h and DlConfig.{h,cc} stay in
// use (none of them needs TensorFlow); funcapproximator/DlTFNetworkEvaluator.cc and DlTensorUtil.cc drop out of the
// build and libugeval.so is linked instead.  tests/test_reference_dropin.py builds search/UctBoardEvaluator.cpp this way.
#ifndef UGEVAL_OVERLAY_DLTFNETWORKEVALUATOR_H
#define UGEVAL_OVERLAY_DLTFNETWORKEVALUATOR_H
#include <string>

#include "config/BoardStaticConfig.h"
#include "funcapproximator/DataFormat.h"
#include "funcapproximator/DlConfig.h"
#define UGEVAL_WITH_REFERENCE_TREE 1
#include "../../DlTFNetworkEvaluator.h"
#endif
