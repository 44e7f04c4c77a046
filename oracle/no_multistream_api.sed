# oracle/no_multistream_api.sed — opus-fix/tests/test_opus_api.c without its multistream sections (SURVEY.md section 2 row 9: the
# multistream wrapper is out of scope) and without the glibc malloc-hook test.  Applied at build time, result in oracle/_ref/gen/.
/^opus_int32 test_msdec_api(void)/,/^}/d
/^int test_malloc_fail(void)/,/^}/d
/total+=test_msdec_api();/d
/total+=test_malloc_fail();/d
/opus_multistream_packet_/d
