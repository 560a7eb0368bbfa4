// Stand-ins for the reference's CONTROL PLANE, test infrastructure only.  search/UctDeepPlayer.cpp (the self-play
// player) links checkpoint download (funcapproximator/DlCheckPoint.cc: libcurl + zip), object upload
// (msg/MinioStub.cc), UUIDs (lib/SgUUID.cpp: libuuid) and three helpers of lib/FileUtil.cc (Boost.Filesystem).  None of
// that is on the evaluation path and none of it is reached by the local self-play loop oracle/_ref/ref_selfplay
// drives (SelfPlayOneGame's body, DoSearch, WriteTFRecord); the symbols only have to exist.
#include <sys/stat.h>
#include <unistd.h>

#include <string>

#include "funcapproximator/DlCheckPoint.h"
#include "lib/FileUtil.h"
#include "lib/SgUUID.h"
#include "msg/MinioStub.h"

DlCheckPoint::CheckPointInfo::CheckPointInfo() {}
bool DlCheckPoint::DownloadMetagraph() { return false; }
bool DlCheckPoint::DownloadCheckPoint(const std::string&, CheckPointInfo&) { return false; }
bool DlCheckPoint::CheckDefaultCheckPoint(CheckPointInfo&) { return false; }
void SgUUID::generateUUID(std::string& out) { out = "00000000-0000-0000-0000-000000000000"; }
void UnrealGo::MinioStub::Upload(const std::string&, const std::string&, const std::string&) {}
std::string UnrealGo::GetFullPathStr(const std::string& directory, const std::string& name) {
  return directory.empty() ? name : directory + "/" + name;
}
std::string UnrealGo::GetCWD() {
  char buf[4096];
  return getcwd(buf, sizeof(buf)) ? std::string(buf) : std::string();
}
int UnrealGo::CreatePath(mode_t mode, const std::string& rootPath, std::string path) {
  const std::string full = rootPath.empty() ? path : rootPath + "/" + path;
  for (size_t i = 1; i <= full.size(); ++i)
    if (i == full.size() || full[i] == '/') mkdir(full.substr(0, i).c_str(), mode);
  return 0;
}
