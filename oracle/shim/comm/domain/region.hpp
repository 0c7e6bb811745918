// oracle/shim/comm/domain/region.hpp -- TEST INFRASTRUCTURE. libcomm comm::Region<T> restated.
#ifndef ORACLE_SHIM_COMM_REGION_H
#define ORACLE_SHIM_COMM_REGION_H
namespace comm {
    template<typename T>
    struct Region {
        union { struct { T x_low, y_low, z_low; }; T low[3]; };
        union { struct { T x_high, y_high, z_high; }; T high[3]; };
        Region() : x_low(0), y_low(0), z_low(0), x_high(0), y_high(0), z_high(0) {}
        Region(T xl, T yl, T zl, T xh, T yh, T zh) : x_low(xl), y_low(yl), z_low(zl), x_high(xh), y_high(yh), z_high(zh) {}
        inline bool isIn(const T x, const T y, const T z) const {
            return x >= x_low && x < x_high && y >= y_low && y < y_high && z >= z_low && z < z_high;
        }
        inline T volume() const { return (x_high - x_low) * (y_high - y_low) * (z_high - z_low); }
    };
}
#endif
