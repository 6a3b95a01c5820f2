// Basic id types of the Wukong surface (reference core/type.hpp:28-127, store/vertex.hpp:33-42).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace wukong {

typedef uint32_t sid_t;   // string id (DTYPE_64BIT off)
typedef int32_t ssid_t;   // signed: negative ids are variables
static const sid_t BLANK_ID = UINT32_MAX;

enum dir_t { IN = 0, OUT = 1, CORUN = 2 };
enum { PREDICATE_ID = 0, TYPE_ID = 1 };
enum { NBITS_IDX = 17 };

static inline bool is_tpid(ssid_t id) { return id > 1 && id < (1 << NBITS_IDX); }
static inline bool is_vid(ssid_t id) { return id >= (1 << NBITS_IDX); }

struct triple_t {
    sid_t s, p, o;
};

enum { SID_t = 0, INT_t = 1, FLOAT_t = 2, DOUBLE_t = 3 };   // utils/variant.hpp

}  // namespace wukong
