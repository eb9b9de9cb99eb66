// so that the reference's example.cpp (#include "pbicgstab.h") picks up the mirror header
#pragma once
#include "../pbicgstab.h"
