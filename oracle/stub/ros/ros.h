// TEST INFRASTRUCTURE ONLY (oracle/): logging-only stand-in for <ros/ros.h> so the reference's
// planner sources compile unmodified outside a ROS workspace.  Nothing here is shipped.
//
// The real ros.h drags in the libstdc++ <math.h>/<stdlib.h> wrappers, which put the floating-point
// overloads of abs() into the global namespace.  The reference calls unqualified abs() on doubles
// (grid_planner_core.cpp:527, bot_utils.cpp:46-47), so the stub must do the same or sparsifyPath
// silently degenerates to int abs().
#ifndef NSFS_ORACLE_STUB_ROS_H
#define NSFS_ORACLE_STUB_ROS_H
#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>
#include <iostream>
#include <sstream>

#define NSFS_STUB_NOP(...) do { } while (0)
#define NSFS_STUB_COND(c, ...) do { (void)sizeof(c); } while (0)

#define ROS_DEBUG(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_INFO(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_WARN(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_ERROR(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_FATAL(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_DEBUG_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_INFO_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_WARN_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_ERROR_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_DEBUG_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_INFO_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_WARN_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_ERROR_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_INFO_STREAM_COND(c, x) NSFS_STUB_COND(c, x)
#define ROS_WARN_STREAM_COND(c, x) NSFS_STUB_COND(c, x)

// ---- just enough of the node API for src/loco_mapping/src/occupancy_grid.cpp to compile and RUN its scan loop ----------
// Parameters come from a global table (nsfs_stub::params), ros::spinOnce() calls a hook the harness installs (it plays the
// role of the pose / scan callbacks), publishers swallow their messages.
#include <functional>
#include <map>
#include <memory>
#include <type_traits>
namespace nsfs_stub {
inline std::map<std::string, double>& params() { static std::map<std::string, double> p; return p; }
inline std::function<void()>& spin_hook() { static std::function<void()> h; return h; }
inline std::function<bool()>& keep_running() { static std::function<bool()> h; return h; }
}
namespace ros {
struct Subscriber {};
struct ServiceServer {};
struct Time { double t = 0; static Time now() { return Time(); } double toSec() const { return t; } };
struct Publisher { template <class M> void publish(const M&) const {} };
struct Rate { explicit Rate(double) {} void sleep() {} };
inline bool ok() { return true; }
inline void spinOnce() { if (nsfs_stub::spin_hook()) nsfs_stub::spin_hook()(); }
class NodeHandle {
public:
    template <class T> bool param(const std::string& name, T& out, const T& def) const {
        if constexpr (std::is_arithmetic<T>::value) {
            auto it = nsfs_stub::params().find(name);
            if (it == nsfs_stub::params().end()) { out = def; return false; }
            out = static_cast<T>(it->second);
            return true;
        } else {                                                      // string parameters (commander.cpp:554,583): always the default
            out = def;
            return false;
        }
    }
    bool param(const std::string& name, bool def) const {          // nh_.param("trigger_nodes", true): the loop condition
        if (name == "trigger_nodes" && nsfs_stub::keep_running()) return nsfs_stub::keep_running()();
        return def;
    }
    template <class M, class C> Subscriber subscribe(const std::string&, int, void (C::*)(M), C*) { return Subscriber(); }
    template <class M> Publisher advertise(const std::string&, int, bool = false) { return Publisher(); }
    // commander.cpp: const-reference callbacks, a service, string / bool parameters with literal defaults
    template <class M, class C> Subscriber subscribe(const std::string&, int, void (C::*)(const M&), C*) { return Subscriber(); }
    template <class Rq, class Rs, class C> ServiceServer advertiseService(const std::string&, bool (C::*)(Rq&, Rs&), C*) { return ServiceServer(); }
};
}

#endif
