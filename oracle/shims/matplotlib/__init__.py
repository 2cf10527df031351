"""TEST INFRASTRUCTURE ONLY. Stub so the unmodified reference imports without matplotlib
(absent in this image). Any plotting call raises."""
