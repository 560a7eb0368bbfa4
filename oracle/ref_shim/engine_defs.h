// Force-included (-include) when the reference's engine sources are compiled for oracle/_ref/ref_search: the string
// definitions its CMakeLists.txt passes on the command line (:6-7, :21), plus <string>, which the reference's
// DlConfig.h uses without including.  Test infrastructure only.
#pragma once
#include <string>
#define APP_NAME "unrealgo"
#define APP_VERSION "1.0"
#define HOST_CPU "x86_64"
