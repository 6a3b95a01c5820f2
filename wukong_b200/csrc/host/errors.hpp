// Status codes and the exception type of the Wukong surface (reference utils/errors.hpp:25-79).
#pragma once
#include <exception>

namespace wukong {

enum {
    SUCCESS = 0, UNKNOWN_ERROR, SYNTAX_ERROR, UNKNOWN_PATTERN, ATTR_DISABLE, NO_REQUIRED_VAR, UNSUPPORT_UNION,
    OBJ_ERROR, VERTEX_INVALID, UNKNOWN_SUB, SETTING_ERROR, FIRST_PATTERN_ERROR, UNKNOWN_FILTER, ERROR_LAST
};

static inline const char *err_msg(int code) {
    static const char *M[ERROR_LAST] = {
        "Everything is ok", "Something wrong happened", "Something wrong in the query syntax, fail to parse!",
        "Unsupported triple pattern.", "MUST enable attribute support!", "NO required variables!",
        "Unsupport UNION on attribute results", "Object should not be an index", "Subject or object is not valid",
        "Triple pattern should not start from unknown subject.", "You may change SETTING files to avoid this error.",
        "Const_X_X or index_X_X must be the first pattern.", "Unsupported filter type."};
    return (code >= 0 && code < ERROR_LAST) ? M[code] : "engine error";
}

struct WukongException : public std::exception {
    int status_code;
    explicit WukongException(int c) : status_code(c) {}
    const char *what() const noexcept override { return err_msg(status_code); }
    int code() const { return status_code; }
};

#define ASSERT_ERROR_CODE(cond, code) do { if (!(cond)) throw ::wukong::WukongException(code); } while (0)

}  // namespace wukong
