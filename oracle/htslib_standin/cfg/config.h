#define VERSION "3.3.0"
