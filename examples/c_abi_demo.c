/*
 * c_abi_demo.c — the C ABI of include/game_engine_b200.h used from plain C (no Python, no torch):
 * load a compiled transition table, create a batch of sessions on GPU 0, step it to the end, print the statistics.
 *
 *   python -c "from game_engine_b200 import compile_game; open('/tmp/w8.getb','wb').write(compile_game('werewolf-(mafia)', 8).blob)"
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o /tmp/ge_demo -Lgame_engine_b200 -lgame_engine_b200 -Wl,-rpath,$PWD/game_engine_b200
 *   /tmp/ge_demo /tmp/w8.getb 1048576 56 7
 *
 * This is what a cgo / JNI / N-API binding of the path would wrap; the reference itself is Python and binds with
 * ctypes.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "game_engine_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != GE_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ge_last_error()); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s table.getb [n_sessions] [n_steps] [seed]\n", argv[0]); return 2; }
    const uint64_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 4096;
    const int steps = argc > 3 ? atoi(argv[3]) : 64;
    const uint64_t seed = argc > 4 ? strtoull(argv[4], NULL, 10) : 1;

    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    static uint8_t blob[8192];
    const size_t len = fread(blob, 1, sizeof blob, f);
    fclose(f);

    ge_table *tab = NULL;
    ge_batch *bat = NULL;
    CHECK(ge_table_create(blob, len, &tab));
    CHECK(ge_batch_create(tab, 0, n, /*first_session_id=*/0, seed, &bat));
    CHECK(ge_step(bat, steps, NULL));                       /* `steps` launches, one session-phase-step each */

    static uint64_t st[GE_STATS_LEN];
    CHECK(ge_stats(bat, st, GE_STATS_LEN));
    const size_t S = ge_table_record_size(tab);
    uint8_t *first = malloc(S);
    CHECK(ge_export_state(bat, 0, 1, first));
    printf("%s\nplayers=%d record=%zuB sessions=%llu steps=%d\n", ge_version(), ge_table_n_players(tab), S,
           (unsigned long long)n, steps);
    printf("counted=%llu winners=[%llu,%llu,%llu] session0: phase_index=%u step=%u\n", (unsigned long long)st[0],
           (unsigned long long)st[1], (unsigned long long)st[2], (unsigned long long)st[3], first[0],
           (unsigned)(first[2] | (first[3] << 8)));
    free(first);
    ge_batch_destroy(bat);
    ge_table_destroy(tab);
    return 0;
}
