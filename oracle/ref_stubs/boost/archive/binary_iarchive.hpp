#pragma once
#include "binary_oarchive.hpp"
