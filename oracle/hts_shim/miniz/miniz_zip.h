#include "miniz.h"
