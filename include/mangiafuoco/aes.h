/* Forwarding header: the reference's aes.h interface lives in mangiafuoco_b200.h. */
#pragma once
#include "mangiafuoco_b200.h"
