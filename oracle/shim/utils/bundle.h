// oracle/shim/utils/bundle.h -- TEST INFRASTRUCTURE. kiwi::Bundle is only named in declarations of
// frontend/config_values.h, which frontend/io/atom_dump.h includes.
#ifndef ORACLE_SHIM_KIWI_BUNDLE_H
#define ORACLE_SHIM_KIWI_BUNDLE_H
namespace kiwi { class Bundle; }
#endif
