/* C3 stand-in: text (de)serialisation of a double, used only by the dead
 * HashGrid code in util.c:352-656. */
#ifndef C3SHIM_STRINGMANIP_H
#define C3SHIM_STRINGMANIP_H
char  *serialize_double_to_text(double v);
double deserialize_double_from_text(char *s);
#endif
