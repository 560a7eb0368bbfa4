// Stand-in for <curl/curl.h> (libcurl is not installed here): network/AtomCurl.h, included by search/UctDeepPlayer.cpp,
// needs this one constant for a member array.  Test infrastructure only.
#pragma once
#define CURL_ERROR_SIZE 256
