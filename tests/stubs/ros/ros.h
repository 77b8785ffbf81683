// stub of <ros/ros.h> (tests/stubs/README.md)
#pragma once
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <typeinfo>
#include <vector>
#define ROS_INFO(...) do { if (::ros::stub::verbose()) { std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)
#define ROS_WARN(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_BREAK() do { std::fprintf(stderr, "ROS_BREAK at %s:%d\n", __FILE__, __LINE__); std::abort(); } while (0)
namespace ros {
namespace stub {
inline bool& verbose() { static bool v = false; return v; }
inline std::map<std::string, double>& params() { static std::map<std::string, double> p; return p; }
struct Published { std::string topic, type; size_t count = 0; };
inline std::map<std::string, Published>& topics() { static std::map<std::string, Published> t; return t; }
}  // namespace stub
struct Time {
  double sec = 0;
  static Time now() { static double t = 0; Time r; r.sec = (t += 0.1); return r; }
};
namespace param {
template <typename T>
bool get(const std::string& key, T& out) {
  auto it = stub::params().find(key);
  if (it == stub::params().end()) return false;
  out = static_cast<T>(it->second);
  return true;
}
}  // namespace param
class Publisher {
 public:
  Publisher() = default;
  explicit Publisher(std::string topic) : topic_(std::move(topic)) {}
  template <typename M>
  void publish(const M&) const { auto& t = stub::topics()[topic_]; t.topic = topic_; t.type = typeid(M).name(); ++t.count; }
 private:
  std::string topic_;
};
class NodeHandle {
 public:
  NodeHandle() = default;
  explicit NodeHandle(const std::string&) {}
  template <typename M>
  Publisher advertise(const std::string& topic, int) { stub::topics()[topic].topic = topic; return Publisher(topic); }
};
}  // namespace ros
