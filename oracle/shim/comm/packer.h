// oracle/shim/comm/packer.h -- TEST INFRASTRUCTURE. libcomm comm::Packer<T> restated (the interface every
// reference packer in src/pack implements).
#ifndef ORACLE_SHIM_COMM_PACKER_H
#define ORACLE_SHIM_COMM_PACKER_H
namespace comm {
    template<typename T>
    class Packer {
    public:
        typedef T pack_date_type;
        virtual ~Packer() {}
        virtual const unsigned long sendLength(const int dimension, const int direction) = 0;
        virtual void onSend(T buffer[], const unsigned long send_len, const int dimension, const int direction) = 0;
        virtual void onReceive(T buffer[], const unsigned long receive_len, const int dimension, const int direction) = 0;
    };
}
#endif
