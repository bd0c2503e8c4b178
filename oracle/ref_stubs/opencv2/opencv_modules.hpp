/* empty stub so the reference's vendored OpenCV headers parse without OpenCV's generated build tree (oracle/Makefile) */
