#pragma once
#include "binary_oarchive.hpp"
namespace boost { namespace archive { typedef binary_oarchive text_oarchive; typedef binary_iarchive text_iarchive; } }
