// openpbso drop-in: compile-time constants.  Mirrors reference config.h:11-14.
#ifndef CONFIG_H
#define CONFIG_H
#define FILE_NOT_EXIST "__NA_FILE"
const static int SAMPLE_RATE = 44100.;
const static int FRAMES_PER_BUFFER = 513;
#endif
