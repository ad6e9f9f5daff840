/* tests/dropin/replay_rand.c — TEST INFRASTRUCTURE: rand() that replays a caller-supplied integer
 * stream (compiled WITHOUT the -Drand= redirection; falls back to libc rand when no stream is set). */
#include <stdlib.h>

static const int *stream;
static size_t stream_len, stream_pos, stream_underflow;

void oracle_set_replay(const int *buf, size_t n) { stream = buf; stream_len = n; stream_pos = 0; stream_underflow = 0; }
size_t oracle_replay_pos(void) { return stream_pos; }
size_t oracle_replay_underflow(void) { return stream_underflow; }
void oracle_srand(unsigned s) { if (!stream) srand(s); }

int replay_rand(void)
{
    if (!stream) return rand();
    if (stream_pos < stream_len) return stream[stream_pos++];
    stream_underflow++;
    return 0;
}

void replay_srand(unsigned s) { oracle_srand(s); }
