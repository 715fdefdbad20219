#pragma once
#include "deplex/config.h"
#include "deplex/plane_extractor.h"
#include "deplex/sequence_extractor.h"
#include "deplex/utils/utils.h"
