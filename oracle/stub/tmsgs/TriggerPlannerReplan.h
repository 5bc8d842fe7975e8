// TEST INFRASTRUCTURE ONLY (oracle/): tmsgs/TriggerPlannerReplan.srv (reference src/tmsgs/srv/TriggerPlannerReplan.srv).
#ifndef NSFS_ORACLE_STUB_TRIGGERPLANNERREPLAN_H
#define NSFS_ORACLE_STUB_TRIGGERPLANNERREPLAN_H
namespace tmsgs { struct TriggerPlannerReplan { struct Request { bool request = false; }; struct Response { bool response = false; }; }; }
#endif
