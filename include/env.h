/* Thin compatibility header: the reference's include/env.h is replaced by the single boundary
 * header ppo_b200.h, which declares the same types and functions (see the citations there). */
#ifndef PPO_B200_COMPAT_ENV_H
#define PPO_B200_COMPAT_ENV_H
#include "ppo_b200.h"
#endif
