// written by oracle/build_ref.py (stands in for include/embree3/rtcore_config.h)
#pragma once
#define RTC_VERSION_MAJOR 3
#define RTC_VERSION_MINOR 12
#define RTC_VERSION_PATCH 1
#define RTC_VERSION 31201
#define RTC_VERSION_STRING "3.12.1"
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1
#define EMBREE_MIN_WIDTH 0
#define RTC_MIN_WIDTH EMBREE_MIN_WIDTH
#define RTC_NAMESPACE_BEGIN
#define RTC_NAMESPACE_END
#define RTC_NAMESPACE_USE
#if defined(__cplusplus)
#  define RTC_API_EXTERN_C extern "C"
#else
#  define RTC_API_EXTERN_C
#endif
#define RTC_API_IMPORT RTC_API_EXTERN_C
#define RTC_API_EXPORT RTC_API_EXTERN_C __attribute__ ((visibility ("default")))
#if defined(RTC_EXPORT_API)
#  define RTC_API RTC_API_EXPORT
#else
#  define RTC_API RTC_API_IMPORT
#endif
