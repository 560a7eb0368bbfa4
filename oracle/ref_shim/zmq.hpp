// Stand-in for <zmq.hpp> (ZeroMQ is not installed here): search/UctDeepPlayer.{h,cpp} include it for the client/server
// self-play modes; the local self-play path that oracle/_ref/ref_selfplay drives never opens a socket.
// Test infrastructure only.
#pragma once
#include <cstddef>
#include <string>
namespace zmq {
class context_t { public: explicit context_t(int = 1) {} };
class message_t {
 public:
  message_t() {}
  explicit message_t(size_t) {}
  message_t(const void*, size_t) {}
  void* data() { return nullptr; }
  size_t size() const { return 0; }
};
class socket_t {
 public:
  socket_t(context_t&, int) {}
  void connect(const std::string&) {}
  void bind(const std::string&) {}
  bool send(message_t&, int = 0) { return false; }
  bool recv(message_t*, int = 0) { return false; }
  void setsockopt(int, const void*, size_t) {}
  void close() {}
};
}
#define ZMQ_REQ 3
#define ZMQ_REP 4
#define ZMQ_PUB 1
#define ZMQ_SUB 2
#define ZMQ_SUBSCRIBE 6
#define ZMQ_LINGER 17
#define ZMQ_RCVTIMEO 27
#define ZMQ_SNDTIMEO 28
