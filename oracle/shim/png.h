// png.h -- declarations only (ours), so demo/png_2d_scenario.hpp parses; its PNG reader/writer is never called.
#pragma once
#include <cstdio>
#include <csetjmp>
typedef unsigned char png_byte; typedef png_byte* png_bytep; typedef png_bytep* png_bytepp;
typedef struct png_struct_def* png_structp; typedef struct png_info_def* png_infop;
#define PNG_LIBPNG_VER_STRING "shim"
#define PNG_COLOR_TYPE_RGB 2
#define PNG_COLOR_TYPE_PALETTE 3
#define PNG_COLOR_MASK_ALPHA 4
#define PNG_INTERLACE_NONE 0
#define PNG_COMPRESSION_TYPE_DEFAULT 0
#define PNG_FILTER_TYPE_DEFAULT 0
png_structp png_create_write_struct(const char*, void*, void*, void*); png_structp png_create_read_struct(const char*, void*, void*, void*);
png_infop png_create_info_struct(png_structp); jmp_buf& png_jmpbuf(png_structp);
void png_init_io(png_structp, FILE*); void png_set_IHDR(png_structp, png_infop, int, int, int, int, int, int, int);
void png_write_info(png_structp, png_infop); void png_write_image(png_structp, png_bytepp); void png_write_end(png_structp, png_infop);
void png_read_info(png_structp, png_infop); png_byte png_get_color_type(png_structp, png_infop); png_byte png_get_bit_depth(png_structp, png_infop);
void png_set_palette_to_rgb(png_structp); void png_set_strip_16(png_structp); void png_set_packing(png_structp); void png_set_strip_alpha(png_structp);
void png_read_update_info(png_structp, png_infop); int png_get_image_width(png_structp, png_infop); int png_get_image_height(png_structp, png_infop);
int png_get_rowbytes(png_structp, png_infop); void png_read_image(png_structp, png_bytepp);
void png_destroy_read_struct(png_structp*, png_infop*, png_infop*); void png_destroy_write_struct(png_structp*, png_infop*);
