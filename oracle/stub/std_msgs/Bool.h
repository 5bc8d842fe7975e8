// TEST INFRASTRUCTURE ONLY (oracle/).
#ifndef NSFS_ORACLE_STUB_BOOL_H
#define NSFS_ORACLE_STUB_BOOL_H
namespace std_msgs { struct Bool { bool data = false; }; }
#endif
