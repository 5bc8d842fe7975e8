// TEST INFRASTRUCTURE ONLY (oracle/): tmsgs/TurtlePath.msg = { nav_msgs/Path path; uint64 id } (reference src/tmsgs/msg/TurtlePath.msg).
#ifndef NSFS_ORACLE_STUB_TURTLEPATH_H
#define NSFS_ORACLE_STUB_TURTLEPATH_H
#include <cstdint>
#include "nav_msgs/Path.h"
namespace tmsgs { struct TurtlePath { nav_msgs::Path path; uint64_t id = 0; }; }
#endif
