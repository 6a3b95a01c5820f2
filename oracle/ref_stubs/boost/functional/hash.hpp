#pragma once
#include "../functional_hash_stub.hpp"
