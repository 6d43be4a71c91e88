// oracle/refcheck/stubs (see Event/Event.h)
#pragma once
#define PROFILE_FUNC()
