/* empty stand-in for the Windows-only <conio.h> the reference includes (pbicgstab.h:17, pbicgstab.cu:19); test-only */
