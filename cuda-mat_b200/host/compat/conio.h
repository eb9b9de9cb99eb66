// conio.h shim: the reference includes this Windows-only header (pbicgstab.h:17, example.cpp:15)
// but uses nothing from it.
#pragma once
