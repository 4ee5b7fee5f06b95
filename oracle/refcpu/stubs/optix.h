#pragma once
#include <optix_stubs.h>
