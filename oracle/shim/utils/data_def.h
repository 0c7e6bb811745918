// oracle/shim/utils/data_def.h -- TEST INFRASTRUCTURE. kiwi basic types.
#ifndef ORACLE_SHIM_KIWI_DATA_DEF_H
#define ORACLE_SHIM_KIWI_DATA_DEF_H
namespace kiwi { typedef int RID; }
#endif
