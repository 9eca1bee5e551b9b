/*
 * libconfig_lite -- reader for the libconfig text format (see header).
 *
 * Recursive-descent parser over an in-memory copy of the file.  `@include`
 * is resolved relative to the including file's directory first, then as
 * given.  Typing rules follow libconfig: integers without '.', 'e' or 'E'
 * are CONFIG_TYPE_INT (INT64 when suffixed with L or out of int range),
 * everything else numeric is CONFIG_TYPE_FLOAT; the typed lookups do not
 * convert between the two.
 */
#include "libconfig_lite.h"

#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const char *s;      /* text */
    size_t pos, len;
    int line;
    const char *file;   /* for diagnostics and relative includes */
    config_t *cfg;
    int depth;          /* include depth */
} parser_t;

static char *dup_str(const char *s, size_t n)
{
    char *r = (char *)malloc(n + 1);
    if (!r) return NULL;
    memcpy(r, s, n);
    r[n] = '\0';
    return r;
}

static config_setting_t *new_setting(int type, char *name, config_setting_t *parent, int line)
{
    config_setting_t *s = (config_setting_t *)calloc(1, sizeof(config_setting_t));
    if (!s) return NULL;
    s->type = type;
    s->name = name;
    s->parent = parent;
    s->line = line;
    return s;
}

static void free_setting(config_setting_t *s)
{
    int i;
    if (!s) return;
    for (i = 0; i < s->nchild; i++) free_setting(s->child[i]);
    free(s->child);
    free(s->name);
    free(s->sval);
    free(s);
}

static int add_child(config_setting_t *parent, config_setting_t *c)
{
    if (parent->nchild == parent->capchild) {
        int ncap = parent->capchild ? 2 * parent->capchild : 8;
        config_setting_t **n = (config_setting_t **)realloc(parent->child, ncap * sizeof(*n));
        if (!n) return 0;
        parent->child = n;
        parent->capchild = ncap;
    }
    parent->child[parent->nchild++] = c;
    return 1;
}

static int fail(parser_t *p, const char *msg)
{
    if (p->cfg->error_text[0] == '\0') {
        snprintf(p->cfg->error_text, sizeof(p->cfg->error_text), "%s", msg);
        snprintf(p->cfg->error_file, sizeof(p->cfg->error_file), "%s", p->file ? p->file : "");
        p->cfg->error_line = p->line;
    }
    return 0;
}

static void skip_ws(parser_t *p)
{
    while (p->pos < p->len) {
        char c = p->s[p->pos];
        if (c == '\n') { p->line++; p->pos++; }
        else if (isspace((unsigned char)c)) p->pos++;
        else if (c == '#') { while (p->pos < p->len && p->s[p->pos] != '\n') p->pos++; }
        else if (c == '/' && p->pos + 1 < p->len && p->s[p->pos + 1] == '/') {
            while (p->pos < p->len && p->s[p->pos] != '\n') p->pos++;
        } else if (c == '/' && p->pos + 1 < p->len && p->s[p->pos + 1] == '*') {
            p->pos += 2;
            while (p->pos + 1 < p->len && !(p->s[p->pos] == '*' && p->s[p->pos + 1] == '/')) {
                if (p->s[p->pos] == '\n') p->line++;
                p->pos++;
            }
            p->pos = (p->pos + 2 <= p->len) ? p->pos + 2 : p->len;
        } else break;
    }
}

static int parse_settings(parser_t *p, config_setting_t *group, int until_brace);
static config_setting_t *parse_value(parser_t *p, char *name, config_setting_t *parent);

static char *parse_string_literal(parser_t *p)
{
    /* one or more adjacent "..." literals, concatenated */
    size_t cap = 64, n = 0;
    char *out = (char *)malloc(cap);
    if (!out) return NULL;
    for (;;) {
        if (p->pos >= p->len || p->s[p->pos] != '"') break;
        p->pos++;
        while (p->pos < p->len && p->s[p->pos] != '"') {
            char c = p->s[p->pos++];
            if (c == '\\' && p->pos < p->len) {
                char e = p->s[p->pos++];
                switch (e) {
                    case 'n': c = '\n'; break;
                    case 't': c = '\t'; break;
                    case 'r': c = '\r'; break;
                    case 'f': c = '\f'; break;
                    case '\\': c = '\\'; break;
                    case '"': c = '"'; break;
                    case 'x': {
                        char hex[3] = {0, 0, 0};
                        if (p->pos + 1 < p->len) { hex[0] = p->s[p->pos]; hex[1] = p->s[p->pos + 1]; p->pos += 2; }
                        c = (char)strtol(hex, NULL, 16);
                        break;
                    }
                    default: c = e; break;
                }
            } else if (c == '\n') p->line++;
            if (n + 2 > cap) {
                char *t;
                cap *= 2;
                t = (char *)realloc(out, cap);
                if (!t) { free(out); return NULL; }
                out = t;
            }
            out[n++] = c;
        }
        if (p->pos >= p->len) { free(out); fail(p, "unterminated string"); return NULL; }
        p->pos++; /* closing quote */
        skip_ws(p);
    }
    out[n] = '\0';
    return out;
}

static config_setting_t *parse_scalar(parser_t *p, char *name, config_setting_t *parent)
{
    const char *s = p->s + p->pos;
    size_t rem = p->len - p->pos;
    config_setting_t *st;

    if (rem >= 1 && s[0] == '"') {
        char *str = parse_string_literal(p);
        if (!str) return NULL;
        st = new_setting(CONFIG_TYPE_STRING, name, parent, p->line);
        if (!st) { free(str); return NULL; }
        st->sval = str;
        return st;
    }
    if (rem >= 4 && strncasecmp(s, "true", 4) == 0 && !(rem > 4 && (isalnum((unsigned char)s[4]) || s[4] == '_'))) {
        p->pos += 4;
        st = new_setting(CONFIG_TYPE_BOOL, name, parent, p->line);
        if (st) st->ival = 1;
        return st;
    }
    if (rem >= 5 && strncasecmp(s, "false", 5) == 0 && !(rem > 5 && (isalnum((unsigned char)s[5]) || s[5] == '_'))) {
        p->pos += 5;
        st = new_setting(CONFIG_TYPE_BOOL, name, parent, p->line);
        if (st) st->ival = 0;
        return st;
    }
    /* number */
    {
        size_t i = 0;
        int is_float = 0, is_hex = 0;
        char buf[128];
        if (i < rem && (s[i] == '+' || s[i] == '-')) i++;
        if (i + 1 < rem && s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X')) {
            is_hex = 1;
            i += 2;
            while (i < rem && isxdigit((unsigned char)s[i])) i++;
        } else {
            size_t start = i;
            while (i < rem && isdigit((unsigned char)s[i])) i++;
            if (i < rem && s[i] == '.') { is_float = 1; i++; while (i < rem && isdigit((unsigned char)s[i])) i++; }
            if (i == start || (i == start + 1 && s[start] == '.')) { fail(p, "syntax error: value expected"); return NULL; }
            if (i < rem && (s[i] == 'e' || s[i] == 'E')) {
                size_t j = i + 1;
                if (j < rem && (s[j] == '+' || s[j] == '-')) j++;
                if (j < rem && isdigit((unsigned char)s[j])) {
                    is_float = 1;
                    while (j < rem && isdigit((unsigned char)s[j])) j++;
                    i = j;
                }
            }
        }
        if (i == 0 || i >= sizeof(buf)) { fail(p, "syntax error: bad number"); return NULL; }
        memcpy(buf, s, i);
        buf[i] = '\0';
        p->pos += i;
        if (is_float) {
            st = new_setting(CONFIG_TYPE_FLOAT, name, parent, p->line);
            if (st) st->fval = strtod(buf, NULL);
            return st;
        } else {
            int is64 = 0;
            long long v;
            errno = 0;
            v = is_hex ? (long long)strtoull(buf, NULL, 16) : strtoll(buf, NULL, 10);
            if (p->pos < p->len && (p->s[p->pos] == 'L' || p->s[p->pos] == 'l')) {
                is64 = 1;
                p->pos++;
                if (p->pos < p->len && (p->s[p->pos] == 'L' || p->s[p->pos] == 'l')) p->pos++;
            }
            if (v > INT_MAX || v < INT_MIN) is64 = 1;
            st = new_setting(is64 ? CONFIG_TYPE_INT64 : CONFIG_TYPE_INT, name, parent, p->line);
            if (st) st->ival = v;
            return st;
        }
    }
}

static config_setting_t *parse_value(parser_t *p, char *name, config_setting_t *parent)
{
    config_setting_t *st;
    skip_ws(p);
    if (p->pos >= p->len) { fail(p, "unexpected end of input"); free(name); return NULL; }
    if (p->s[p->pos] == '{') {
        p->pos++;
        st = new_setting(CONFIG_TYPE_GROUP, name, parent, p->line);
        if (!st) { free(name); return NULL; }
        if (!parse_settings(p, st, 1)) { free_setting(st); return NULL; }
        return st;
    }
    if (p->s[p->pos] == '(' || p->s[p->pos] == '[') {
        char close = (p->s[p->pos] == '(') ? ')' : ']';
        int type = (close == ')') ? CONFIG_TYPE_LIST : CONFIG_TYPE_ARRAY;
        p->pos++;
        st = new_setting(type, name, parent, p->line);
        if (!st) { free(name); return NULL; }
        for (;;) {
            config_setting_t *el;
            skip_ws(p);
            if (p->pos >= p->len) { fail(p, "unterminated list"); free_setting(st); return NULL; }
            if (p->s[p->pos] == close) { p->pos++; break; }
            el = parse_value(p, NULL, st);
            if (!el) { free_setting(st); return NULL; }
            if (!add_child(st, el)) { free_setting(el); free_setting(st); return NULL; }
            skip_ws(p);
            if (p->pos < p->len && p->s[p->pos] == ',') p->pos++;
        }
        return st;
    }
    st = parse_scalar(p, name, parent);
    if (!st) free(name);
    return st;
}

static char *read_file(const char *path, size_t *len)
{
    FILE *f = fopen(path, "rb");
    char *buf;
    long sz;
    if (!f) return NULL;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return NULL; }
    sz = ftell(f);
    if (sz < 0) { fclose(f); return NULL; }
    rewind(f);
    buf = (char *)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); return NULL; }
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { free(buf); fclose(f); return NULL; }
    fclose(f);
    buf[sz] = '\0';
    *len = (size_t)sz;
    return buf;
}

static int parse_include(parser_t *p, config_setting_t *group)
{
    char *fname, *text = NULL;
    char path[1024];
    size_t len = 0;
    parser_t sub;
    int ok;

    skip_ws(p);
    if (p->pos >= p->len || p->s[p->pos] != '"') return fail(p, "@include expects a quoted file name");
    fname = parse_string_literal(p);
    if (!fname) return 0;
    if (p->depth >= 10) { free(fname); return fail(p, "@include nesting too deep"); }

    path[0] = '\0';
    if (fname[0] != '/' && p->file) {
        const char *slash = strrchr(p->file, '/');
        if (slash) {
            size_t dl = (size_t)(slash - p->file) + 1;
            if (dl + strlen(fname) < sizeof(path)) {
                memcpy(path, p->file, dl);
                strcpy(path + dl, fname);
                text = read_file(path, &len);
            }
        }
    }
    if (!text) {
        snprintf(path, sizeof(path), "%s", fname);
        text = read_file(path, &len);
    }
    free(fname);
    if (!text) return fail(p, "cannot open include file");

    sub.s = text; sub.pos = 0; sub.len = len; sub.line = 1;
    sub.file = path; sub.cfg = p->cfg; sub.depth = p->depth + 1;
    ok = parse_settings(&sub, group, 0);
    free(text);
    return ok;
}

static int parse_settings(parser_t *p, config_setting_t *group, int until_brace)
{
    for (;;) {
        size_t start;
        char *name;
        config_setting_t *val;

        skip_ws(p);
        if (p->pos >= p->len) {
            if (until_brace) return fail(p, "missing '}'");
            return 1;
        }
        if (p->s[p->pos] == '}') {
            if (!until_brace) return fail(p, "unexpected '}'");
            p->pos++;
            return 1;
        }
        if (p->s[p->pos] == '@') {
            if (p->len - p->pos >= 8 && strncmp(p->s + p->pos, "@include", 8) == 0) {
                p->pos += 8;
                if (!parse_include(p, group)) return 0;
                continue;
            }
            return fail(p, "unknown directive");
        }
        start = p->pos;
        if (!(isalpha((unsigned char)p->s[p->pos]) || p->s[p->pos] == '*' || p->s[p->pos] == '_'))
            return fail(p, "syntax error: setting name expected");
        while (p->pos < p->len && (isalnum((unsigned char)p->s[p->pos]) || p->s[p->pos] == '_' ||
                                   p->s[p->pos] == '-' || p->s[p->pos] == '*'))
            p->pos++;
        name = dup_str(p->s + start, p->pos - start);
        if (!name) return fail(p, "out of memory");
        skip_ws(p);
        if (p->pos >= p->len || (p->s[p->pos] != '=' && p->s[p->pos] != ':')) {
            free(name);
            return fail(p, "syntax error: '=' or ':' expected");
        }
        p->pos++;
        if (config_setting_get_member(group, name)) {
            free(name);
            return fail(p, "duplicate setting name");
        }
        val = parse_value(p, name, group); /* takes ownership of name */
        if (!val) return 0;
        if (!add_child(group, val)) { free_setting(val); return fail(p, "out of memory"); }
        skip_ws(p);
        if (p->pos < p->len && (p->s[p->pos] == ';' || p->s[p->pos] == ',')) p->pos++;
    }
}

void config_init(config_t *config)
{
    memset(config, 0, sizeof(*config));
    config->root = new_setting(CONFIG_TYPE_GROUP, NULL, NULL, 0);
}

void config_destroy(config_t *config)
{
    if (!config) return;
    free_setting(config->root);
    config->root = NULL;
}

static int parse_text(config_t *config, const char *text, size_t len, const char *file)
{
    parser_t p;
    if (!config->root) config_init(config);
    config->error_text[0] = '\0';
    p.s = text; p.pos = 0; p.len = len; p.line = 1; p.file = file; p.cfg = config; p.depth = 0;
    return parse_settings(&p, config->root, 0) ? CONFIG_TRUE : CONFIG_FALSE;
}

int config_read_file(config_t *config, const char *filename)
{
    size_t len = 0;
    char *text = read_file(filename, &len);
    int ok;
    if (!text) {
        snprintf(config->error_text, sizeof(config->error_text), "file I/O error");
        snprintf(config->error_file, sizeof(config->error_file), "%s", filename);
        return CONFIG_FALSE;
    }
    ok = parse_text(config, text, len, filename);
    free(text);
    return ok;
}

int config_read_string(config_t *config, const char *text)
{
    return parse_text(config, text, strlen(text), NULL);
}

config_setting_t *config_root_setting(const config_t *config) { return config->root; }

int config_setting_type(const config_setting_t *s) { return s ? s->type : CONFIG_TYPE_NONE; }

const char *config_setting_name(const config_setting_t *s) { return s ? s->name : NULL; }

static int is_aggregate(const config_setting_t *s)
{
    return s && (s->type == CONFIG_TYPE_GROUP || s->type == CONFIG_TYPE_LIST || s->type == CONFIG_TYPE_ARRAY);
}

int config_setting_length(const config_setting_t *s) { return is_aggregate(s) ? s->nchild : 0; }

config_setting_t *config_setting_get_elem(const config_setting_t *s, unsigned int idx)
{
    if (!is_aggregate(s) || idx >= (unsigned int)s->nchild) return NULL;
    return s->child[idx];
}

config_setting_t *config_setting_get_member(const config_setting_t *s, const char *name)
{
    int i;
    if (!s || s->type != CONFIG_TYPE_GROUP || !name) return NULL;
    for (i = 0; i < s->nchild; i++)
        if (s->child[i]->name && strcmp(s->child[i]->name, name) == 0) return s->child[i];
    return NULL;
}

config_setting_t *config_lookup(const config_t *config, const char *path)
{
    /* path components separated by '.', ':' or '/'; "[n]" indexes lists */
    config_setting_t *cur = config->root;
    const char *q = path;
    while (cur && *q) {
        char comp[256];
        size_t n = 0;
        while (*q == '.' || *q == ':' || *q == '/') q++;
        if (!*q) break;
        if (*q == '[') {
            long idx = strtol(q + 1, (char **)&q, 10);
            if (*q == ']') q++;
            cur = config_setting_get_elem(cur, (unsigned int)idx);
            continue;
        }
        while (*q && *q != '.' && *q != ':' && *q != '/' && *q != '[' && n + 1 < sizeof(comp)) comp[n++] = *q++;
        comp[n] = '\0';
        cur = config_setting_get_member(cur, comp);
    }
    return cur;
}

int config_setting_lookup_int(const config_setting_t *s, const char *name, int *value)
{
    config_setting_t *m = config_setting_get_member(s, name);
    if (!m || m->type != CONFIG_TYPE_INT) return CONFIG_FALSE;
    *value = (int)m->ival;
    return CONFIG_TRUE;
}

int config_setting_lookup_int64(const config_setting_t *s, const char *name, long long *value)
{
    config_setting_t *m = config_setting_get_member(s, name);
    if (!m || (m->type != CONFIG_TYPE_INT64 && m->type != CONFIG_TYPE_INT)) return CONFIG_FALSE;
    *value = m->ival;
    return CONFIG_TRUE;
}

int config_setting_lookup_float(const config_setting_t *s, const char *name, double *value)
{
    config_setting_t *m = config_setting_get_member(s, name);
    if (!m || m->type != CONFIG_TYPE_FLOAT) return CONFIG_FALSE;
    *value = m->fval;
    return CONFIG_TRUE;
}

int config_setting_lookup_bool(const config_setting_t *s, const char *name, int *value)
{
    config_setting_t *m = config_setting_get_member(s, name);
    if (!m || m->type != CONFIG_TYPE_BOOL) return CONFIG_FALSE;
    *value = (int)m->ival;
    return CONFIG_TRUE;
}

int config_setting_lookup_string(const config_setting_t *s, const char *name, const char **value)
{
    config_setting_t *m = config_setting_get_member(s, name);
    if (!m || m->type != CONFIG_TYPE_STRING) return CONFIG_FALSE;
    *value = m->sval;
    return CONFIG_TRUE;
}

const char *config_error_text(const config_t *config) { return config->error_text; }
const char *config_error_file(const config_t *config) { return config->error_file; }
int config_error_line(const config_t *config) { return config->error_line; }
