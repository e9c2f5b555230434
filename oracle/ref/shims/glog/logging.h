// glog/logging.h — minimal functional glog shim (test infrastructure): the
// container has no glog; the reference's CPU sources use LOG/VLOG/CHECK*.
#ifndef ORACLE_REF_SHIM_GLOG_H_
#define ORACLE_REF_SHIM_GLOG_H_
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace shim_glog {
inline int& verbosity() {
  static int v = [] {
    const char* e = std::getenv("GLOG_v");
    return e ? std::atoi(e) : 0;
  }();
  return v;
}
class Message {
 public:
  Message(const char* file, int line, int severity) : sev_(severity) {
    static const char* names[] = {"I", "W", "E", "F"};
    ss_ << names[severity] << " " << file << ":" << line << "] ";
  }
  ~Message() {
    if (sev_ >= 1) std::cerr << ss_.str() << std::endl;
    else if (verbosity() >= 0 && std::getenv("GLOG_logtostderr")) std::cerr << ss_.str() << std::endl;
    if (sev_ == 3) std::abort();
  }
  std::ostream& stream() { return ss_; }

 private:
  std::ostringstream ss_;
  int sev_;
};
struct Voidify {
  void operator&(std::ostream&) {}
};
}  // namespace shim_glog

namespace google {
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
inline void ShutdownGoogleLogging() {}
inline void ShutDownCommandLineFlags() {}
}  // namespace google

#define SHIM_SEV_INFO 0
#define SHIM_SEV_WARNING 1
#define SHIM_SEV_ERROR 2
#define SHIM_SEV_FATAL 3
#define LOG(sev) ::shim_glog::Message(__FILE__, __LINE__, SHIM_SEV_##sev).stream()
#define LOG_IF(sev, cond) !(cond) ? (void) 0 : ::shim_glog::Voidify() & LOG(sev)
#define VLOG_IS_ON(n) (::shim_glog::verbosity() >= (n))
#define VLOG(n) !VLOG_IS_ON(n) ? (void) 0 : ::shim_glog::Voidify() & ::shim_glog::Message(__FILE__, __LINE__, 1).stream()
#define CHECK(cond) (cond) ? (void) 0 : ::shim_glog::Voidify() & LOG(FATAL) << "Check failed: " #cond " "
#define SHIM_CHECK_OP(a, b, op) \
  ((a) op (b)) ? (void) 0 : ::shim_glog::Voidify() & LOG(FATAL) << "Check failed: " #a " " #op " " #b " "
#define CHECK_EQ(a, b) SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) SHIM_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) SHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) SHIM_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) SHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) SHIM_CHECK_OP(a, b, >=)
#define DCHECK(cond) CHECK(cond)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define CHECK_NOTNULL(p) (p)
#define PLOG(sev) LOG(sev)

#endif  // ORACLE_REF_SHIM_GLOG_H_
