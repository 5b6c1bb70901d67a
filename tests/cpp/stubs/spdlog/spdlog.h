#pragma once
#include <string>
namespace spdlog {
template <class... A> void debug(const A&...); template <class... A> void info(const A&...); template <class... A> void warn(const A&...);
template <class... A> void error(const A&...); template <class... A> void critical(const A&...); template <class... A> void trace(const A&...);
}
