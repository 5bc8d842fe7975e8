// TEST INFRASTRUCTURE ONLY (oracle/): tmsgs/Brake.srv (reference src/tmsgs/srv/Brake.srv).
#ifndef NSFS_ORACLE_STUB_BRAKE_H
#define NSFS_ORACLE_STUB_BRAKE_H
#include <cstdint>
namespace tmsgs { struct Brake { struct Request { uint8_t brake_mode = 0; }; struct Response { bool response = false; }; }; }
#endif
