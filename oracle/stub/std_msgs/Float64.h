// TEST INFRASTRUCTURE ONLY (oracle/).
#ifndef NSFS_ORACLE_STUB_FLOAT64_H
#define NSFS_ORACLE_STUB_FLOAT64_H
#include <memory>
namespace std_msgs {
struct Float64 { double data = 0; };
typedef std::shared_ptr<const Float64> Float64ConstPtr;
}
#endif
