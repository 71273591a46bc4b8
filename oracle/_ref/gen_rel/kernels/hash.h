#define RTC_HASH "oracle_build_ref"
