#pragma once
#include "common.hpp"
