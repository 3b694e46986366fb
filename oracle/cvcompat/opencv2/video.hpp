#pragma once
#include "video/background_segm.hpp"
