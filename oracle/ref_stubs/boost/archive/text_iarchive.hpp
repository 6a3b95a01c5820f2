#pragma once
#include "text_oarchive.hpp"
