/* oracle/shim/oracle_rand.c — TEST INFRASTRUCTURE (oracle build only).
 *
 * The reference draws every random number from libc rand() (SMC.c:290,335;
 * matematicose.c:188-189).  The reference build is compiled with
 * -Drand=oracle_rand -Dsrand=oracle_srand so that a test can REPLAY a known
 * integer stream into it (the same stream is turned into host-fed Gaussians /
 * uniforms for the CUDA path).  Without a replay buffer it is plain glibc
 * rand()/srand().  This file itself is compiled WITHOUT those -D flags.
 */
#include <stdlib.h>
#include <stdio.h>

static const int *g_replay = NULL;
static size_t g_len = 0, g_pos = 0, g_underflow = 0;

void oracle_set_replay(const int *buf, size_t n) { g_replay = buf; g_len = n; g_pos = 0; g_underflow = 0; }
size_t oracle_replay_pos(void) { return g_pos; }
size_t oracle_replay_underflow(void) { return g_underflow; }

int oracle_rand(void)
{
    if (g_replay) {
        if (g_pos < g_len) return g_replay[g_pos++];
        g_underflow++;
        return 0;
    }
    return rand();
}

/* With a replay buffer installed srand is ignored (sMC calls srand(time(NULL)),
 * SMC.c:40, which must not disturb a replayed run). */
void oracle_srand(unsigned s) { if (!g_replay) srand(s); }

int oracle_rand_max(void) { return RAND_MAX; }
