// Stand-in for <boost/filesystem.hpp>: lib/FileUtil.h only *declares* functions over fs::path, and the
// reference TUs compiled into oracle/_ref never call them.  Test infrastructure only.
#pragma once
#include <string>
#include <vector>
namespace boost { namespace filesystem { class path { public: path() {} path(const std::string&) {} }; } }
