/*
 * libconfig_lite -- a small reader for the libconfig text format.
 *
 * Host-side part of the b200sph drop-in: the reference parses material.cfg with
 * the external libconfig library (reference: src/io.cu:64-73,
 * src/config_parameter.cu:357-878).  That library is not available in this
 * image, so this file provides the subset of its C API that the material
 * reader needs, with the same typing rules (a float lookup on an integer
 * literal fails and vice versa).
 *
 * Supported grammar: settings `name = value` / `name : value`, optional `;`
 * or `,` terminators, groups `{}`, lists `()`, arrays `[]`, ints (dec / hex,
 * optional L suffix), floats, booleans, strings (adjacent strings are
 * concatenated), comments `#`, `//`, `/ * * /`, and `@include "file"`.
 */
#ifndef B200SPH_LIBCONFIG_LITE_H
#define B200SPH_LIBCONFIG_LITE_H

#ifdef __cplusplus
extern "C" {
#endif

#define CONFIG_TRUE 1
#define CONFIG_FALSE 0

enum {
    CONFIG_TYPE_NONE = 0,
    CONFIG_TYPE_GROUP,
    CONFIG_TYPE_INT,
    CONFIG_TYPE_INT64,
    CONFIG_TYPE_FLOAT,
    CONFIG_TYPE_STRING,
    CONFIG_TYPE_BOOL,
    CONFIG_TYPE_ARRAY,
    CONFIG_TYPE_LIST
};

typedef struct config_setting_t {
    char *name;                       /* NULL for list/array elements */
    int type;
    long long ival;                   /* INT, INT64, BOOL */
    double fval;                      /* FLOAT */
    char *sval;                       /* STRING */
    struct config_setting_t **child;  /* GROUP, LIST, ARRAY */
    int nchild, capchild;
    struct config_setting_t *parent;
    int line;
} config_setting_t;

typedef struct config_t {
    config_setting_t *root;
    char error_text[256];
    char error_file[512];
    int error_line;
} config_t;

void config_init(config_t *config);
void config_destroy(config_t *config);
int config_read_file(config_t *config, const char *filename);
int config_read_string(config_t *config, const char *text);

config_setting_t *config_lookup(const config_t *config, const char *path);
config_setting_t *config_root_setting(const config_t *config);
int config_setting_length(const config_setting_t *setting);
config_setting_t *config_setting_get_elem(const config_setting_t *setting, unsigned int idx);
config_setting_t *config_setting_get_member(const config_setting_t *setting, const char *name);
int config_setting_type(const config_setting_t *setting);
const char *config_setting_name(const config_setting_t *setting);

int config_setting_lookup_int(const config_setting_t *setting, const char *name, int *value);
int config_setting_lookup_int64(const config_setting_t *setting, const char *name, long long *value);
int config_setting_lookup_float(const config_setting_t *setting, const char *name, double *value);
int config_setting_lookup_bool(const config_setting_t *setting, const char *name, int *value);
int config_setting_lookup_string(const config_setting_t *setting, const char *name, const char **value);

const char *config_error_text(const config_t *config);
const char *config_error_file(const config_t *config);
int config_error_line(const config_t *config);

#ifdef __cplusplus
}
#endif
#endif
