// oracle/shim/io/io_writer.h -- TEST INFRASTRUCTURE. kiwi I/O is outside the hot path; nothing to declare.
#ifndef ORACLE_SHIM_KIWI_IO_WRITER_H
#define ORACLE_SHIM_KIWI_IO_WRITER_H
#endif
